#!/bin/bash
# development build of the warp-pair kernel only (headline instantiation), optional extra flags: scripts/pair_build.sh -DBFB_PAIR_TIMING
set -e
cd "$(dirname "$0")/../bayesfast_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DBFB_PAIR_HEADLINE_ONLY "$@" -Xptxas -v -c bfb_sampler_pair.cu -o bfb_sampler_pair.o 2>&1 | grep -E "registers|spill" | head -4
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -c bfb_sampler.cu -o bfb_sampler.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libbfb200.so bfb_model.o bfb_sampler.o bfb_sampler_fast.o bfb_sampler_dmma.o bfb_sampler_team.o bfb_sampler_pair.o bfb_eval_dmma.o bfb_lik_dmma.o bfb_micro.o bfb_fit.o bfb_post.o bfb_sampler_dmma_headline.o -cudart static
