#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): bench lines of every workload, the reference
# arm, ncu launch lists of one drop-in step, `ncu --set full` captures of the dominant kernels,
# in-kernel timelines of the two k-means kernels.
#   gpurun -- 'bash scripts/collect_evidence.sh r2b'     ->  gpurun_out/<tag>_*
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py > $O/${TAG}_bench_voc_scribble_b1.json 2> $O/${TAG}_bench_voc_scribble_b1.err
for W in voc_scribble_b4 voc_tag_b2 densepose_b1 voc_scribble_softmax_b1; do
  timeout 600 python bench.py --no-train-arm --workload $W > $O/${TAG}_bench_$W.json 2> $O/${TAG}_bench_$W.err
done
timeout 900 python bench.py --impl reference > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
for W in voc_scribble_b1 voc_scribble_b4; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${TAG}_launches_$W.csv python scripts/profile_step.py --workload $W --steps 3 > /dev/null 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmeans_small -s 2 -c 1 -f \
  -o $O/${TAG}_full_kmeans_b1 python scripts/profile_step.py --workload voc_scribble_b1 --steps 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_normalize_pack -s 2 -c 1 -f \
  -o $O/${TAG}_full_prepass_b1 python scripts/profile_step.py --workload voc_scribble_b1 --steps 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmeans_cluster -s 2 -c 1 -f \
  -o $O/${TAG}_full_kmeans_b4 python scripts/profile_step.py --workload voc_scribble_b4 --steps 3 > /dev/null 2>&1
timeout 300 python scripts/time_kmeans_paths.py voc_scribble_b1 voc_scribble_b1@2 voc_scribble_b1@3 voc_scribble_b4 voc_scribble_b1@8 voc_scribble_b1@16 voc_tag_b2 voc_tag_b2@4 2>&1 | grep -v "^  fp32" > $O/${TAG}_kmeans_paths.txt
SPML_B200_LIB=spml_b200/libspml_b200_trace.so timeout 300 python scripts/trace_kmeans_small.py voc_scribble_b1 2>&1 | tail -24 | cut -c1-1500 > $O/${TAG}_kmeans_small_timeline_b1.txt
SPML_B200_KMEANS=cluster SPML_B200_LIB=spml_b200/libspml_b200_trace.so timeout 300 python scripts/trace_kmeans_cluster.py voc_scribble_b1 2>&1 | tail -12 > $O/${TAG}_kmeans_cluster_timeline_b1.txt
SPML_B200_LIB=spml_b200/libspml_b200_trace.so timeout 300 python scripts/trace_kmeans_cluster.py voc_scribble_b4 2>&1 | tail -12 > $O/${TAG}_kmeans_cluster_timeline_b4.txt
timeout 300 python scripts/ab_step.py SPML_B200_FUSED_PREPASS 0 1 2>&1 | tail -6 > $O/${TAG}_ab_prepass.txt
for W in voc_scribble_b1 voc_scribble_b4 voc_tag_b2 densepose_b1 voc_scribble_softmax_b1; do
  python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench_$W.json').read().strip().splitlines()[-1])
cb=d.get('cpu_baseline') or {}
print('$W', 'resident %.3f ms, e2e %.3f ms, reference %s ms, static %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], cb.get('ms_per_step'), (d.get('static_head') or {}).get('ms_per_step')))
PY
done
tail -c 300 $O/${TAG}_bench_reference.json; echo
cat $O/${TAG}_kmeans_paths.txt | head -40
