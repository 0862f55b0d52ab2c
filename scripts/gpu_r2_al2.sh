timeout 120 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -x -k "vit_class or vit_attention" 2>&1 | tail -1
for f in 3 7; do VB_ATTN_PP_FLAGS=$f timeout 60 python scripts/bench_attn.py 2>&1 | grep tcgen05: | sed "s/^/flags=$f /"; done
VB_ATTN_PP_FLAGS=3 timeout 60 python scripts/micro/pp_trace.py > gpurun_out/al_trace3.log 2>&1
