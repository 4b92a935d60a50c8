set -x
python bench.py > gpurun_out/final_bench_pems07.json 2> gpurun_out/final_bench_pems07.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --verbose 2> gpurun_out/final_verbose.log | tail -1 > gpurun_out/final_bench_verbose.json
cp gpurun_out/breakdown_pems07_train_b64_tf32x3_n1.json gpurun_out/final_breakdown_pems07.json
rm -f gpurun_out/final_workloads.jsonl
for w in pems04 pems08 pems03 gambia syn4096; do python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/final_workloads.jsonl; done
python bench.py --workload pems08 --mode infer --sweep 32,64,128,256,512,1024 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null >> gpurun_out/final_workloads.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 9000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
tail -2 gpurun_out/final_ncu.log
wc -l gpurun_out/final_launches.csv gpurun_out/final_workloads.jsonl
