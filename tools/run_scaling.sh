#!/bin/bash
# Strong-scaling (BASELINE config 5: Levy-100D, n=8192, 2^26 candidates) and weak-scaling (headline workload) bench
# lines at N GPUs of one box:  gpurun --gpus N -- bash tools/run_scaling.sh N
N=${1:-8}
PORT=$((29600 + N))
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"
if [ "$N" = "1" ]; then RUN="python"; fi
timeout 900 $RUN bench.py --gpus $N --scaling strong --pool-total 67108864 --nobs 8192 --dim 100 --objective levy \
    --steps 5 --warmup 2 --no-fit --no-cpu > gpurun_out/scale_strong_n$N.json 2> gpurun_out/scale_strong_n$N.err
echo "strong N=$N rc=$?"
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 3 --no-fit --no-cpu > gpurun_out/scale_weak_n$N.json 2> gpurun_out/scale_weak_n$N.err
echo "weak N=$N rc=$?"
tail -n 3 gpurun_out/scale_strong_n$N.err
