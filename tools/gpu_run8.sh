set -x
mkdir -p gpurun_out/r02
CB_PROFILE_CREATE=1 timeout 300 python tools/e2e_profile.py > gpurun_out/r02/e2e_profile8.log 2>&1
tail -32 gpurun_out/r02/e2e_profile8.log
CB_PROFILE_PIPELINE=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02/cfg5_n2.json 2> gpurun_out/r02/cfg5_n2.err
grep pipeline gpurun_out/r02/cfg5_n2.json gpurun_out/r02/cfg5_n2.err | tail -16; tail -c 300 gpurun_out/r02/cfg5_n2.json
