# e2e of eight GPUs draining into one host for several report-chunk counts (BFB200_E2E_CHUNKS; default 64)
for ch in ${CHUNKS:-32 64 128}; do
  BFB200_E2E_CHUNKS=$ch python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_n8_chunks$ch.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n8_chunks$ch.json").read())
print("chunks $ch value %.3e ms %.1f e2e %.3e e2e_ms %.1f kernel_ms_in_e2e %.1f d2h %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["kernel_ms_per_step"], d["e2e"]["d2h_bytes_per_step"]))
PY
done
