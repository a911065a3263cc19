mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 300 python scripts/time_kmeans_paths.py voc_scribble_b1 voc_scribble_b4 2>&1 | grep -v "^ *tc\|^ *fp32" | tail -14
SPML_B200_LIB=spml_b200/libspml_b200_trace.so timeout 300 python scripts/trace_kmeans_small.py voc_scribble_b1 2>&1 | tail -24 | cut -c1-1500 > gpurun_out/kms_trace_b1.txt
tail -9 gpurun_out/kms_trace_b1.txt | cut -c1-700
