timeout 180 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -x -k "vit_class or vit_attention" 2>&1 | tail -3
timeout 60 python scripts/bench_attn.py 2>&1 | grep tcgen05:
timeout 60 python scripts/micro/pp_trace.py > gpurun_out/al_trace3.log 2>&1
