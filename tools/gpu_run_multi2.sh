set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02/test_multi2.log; tail -5 gpurun_out/r02/test_multi2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02/bench_n2.json 2> gpurun_out/r02/bench_n2.err; tail -c 1500 gpurun_out/r02/bench_n2.err; head -c 600 gpurun_out/r02/bench_n2.json
