# usage: gpu_multi.sh N
mkdir -p gpurun_out
N=$1
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
for w in c2 c5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 100 --warmup 10 --no-extras > gpurun_out/bench_multi_${w}_$N.json 2> gpurun_out/bench_multi_${w}_$N.err
tail -c 1500 gpurun_out/bench_multi_${w}_$N.json; tail -5 gpurun_out/bench_multi_${w}_$N.err
done
