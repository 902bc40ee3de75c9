set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 10 --warmup 3 --no-extra > gpurun_out/s4e_bench_n$1.json 2> gpurun_out/s4e_bench_n$1.err; echo rc=$?
tail -c 2500 gpurun_out/s4e_bench_n$1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $1 --steps 2 --warmup 1 > gpurun_out/s4e_bench_reference_n$1.json 2> gpurun_out/s4e_bench_reference_n$1.err; echo rc=$?
tail -c 600 gpurun_out/s4e_bench_reference_n$1.json
