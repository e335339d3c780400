N=${1:-2}
mkdir -p gpurun_out/scale
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale/bench_n$N.json 2> gpurun_out/scale/bench_n$N.err
echo "exit $?"; tail -c 1500 gpurun_out/scale/bench_n$N.json; tail -3 gpurun_out/scale/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/scale/ref_n$N.json 2> gpurun_out/scale/ref_n$N.err
echo "ref exit $?"; tail -c 600 gpurun_out/scale/ref_n$N.json
