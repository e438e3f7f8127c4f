N=$1
mkdir -p gpurun_out
(nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA|Core"; nvidia-smi topo -m) > gpurun_out/boxinfo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 benchmarks/dist_check.py > gpurun_out/r02_dist_check_${N}gpu.log 2>&1
tail -5 gpurun_out/r02_dist_check_${N}gpu.log
timeout 700 $TR --master-port 29512 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 600 gpurun_out/r02_bench_n$N.err
timeout 300 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_n$N.json 2> gpurun_out/r02_bench_reference_n$N.err
grep "^{" gpurun_out/r02_bench_reference_n$N.json | cut -c1-400
grep "^{" gpurun_out/r02_bench_n$N.json | cut -c1-300
