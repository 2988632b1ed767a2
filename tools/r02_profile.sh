set -u
mkdir -p gpurun_out
# (a) launch list of a shortened bench run
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --skip-e2e --later-picks 0 --mi-candidates 20000000 --km-rows 300000 > gpurun_out/r02_final_launches_bench.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_final_launches.csv > gpurun_out/r02_final_launches.txt 2>&1
# (b) MI byte-stream kernel, full set
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mi_stream8 -c 1 -o gpurun_out/r02_mi_bytes_final -f python tools/mi_one.py 100000000 1024 4 bytes 20 > gpurun_out/r02_mi_bytes_final.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_mi_bytes_final.ncu-rep > gpurun_out/r02_mi_bytes_final.ncu.txt
python tools/ncu_hot_sass.py gpurun_out/r02_mi_bytes_final.ncu-rep > gpurun_out/r02_mi_bytes_final.hot.txt 2>&1
python tools/ncu_traffic.py gpurun_out/r02_mi_bytes_final.ncu-rep mi_stream8_kernel 4
# (c) k-means distance GEMM, full set
timeout 300 ncu --set full --clock-control none -k regex:km_assign_pair -s 3 -c 1 -o gpurun_out/r02_km_pair256_final -f python tools/km_tile_bench.py 2048 1024 2 131072 > gpurun_out/r02_km_pair256_final.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_km_pair256_final.ncu-rep > gpurun_out/r02_km_pair256_final.ncu.txt
python tools/ncu_traffic.py gpurun_out/r02_km_pair256_final.ncu-rep km_assign_pair_kernel 1
# (d) loader throughput
timeout 400 python tools/cluster_throughput.py > gpurun_out/r02_cluster_throughput.json 2> gpurun_out/r02_cluster_throughput.err
cp profiles/ncu_traffic.json gpurun_out/r02_ncu_traffic.json
tail -3 gpurun_out/r02_final_launches.txt; cat gpurun_out/r02_ncu_traffic.json; tail -c 600 gpurun_out/r02_cluster_throughput.json
