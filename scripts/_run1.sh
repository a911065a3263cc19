mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "kmeans" 2>&1 | tail -3
timeout 300 python scripts/time_kmeans_paths.py voc_scribble_b1 voc_scribble_b1@2 voc_scribble_b4 2>&1 | grep -v "fp32" | tail -12 | tee gpurun_out/kmeans_paths.txt
SPML_B200_KMEANS=cluster SPML_B200_LIB=spml_b200/libspml_b200_trace.so timeout 300 python scripts/trace_kmeans_cluster.py voc_scribble_b1 2>&1 | tail -12 | tee gpurun_out/kc_trace_b1.txt | cut -c1-600
