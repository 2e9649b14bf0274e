# e2e of one GPU for several report-chunk counts (BFB200_E2E_CHUNKS; default 64)
for ch in ${CHUNKS:-16 32 64 128}; do
  BFB200_E2E_CHUNKS=$ch python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_n1_chunks$ch.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_chunks$ch.json").read())
print("chunks $ch value %.3e ms %.2f e2e %.3e e2e_ms %.2f kernel_ms_in_e2e %.2f d2h %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["kernel_ms_per_step"], d["e2e"]["d2h_bytes_per_step"]))
PY
done
