#!/bin/bash
# Round-2 multi-GPU session (gpurun --gpus 8): concurrent PCIe ceiling, one-process sharded solve, two-device
# shard equivalence, and the 8-rank bench line with the secondary configs (config 5: global 2^23 sharded).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
python tools/pcie_probe.py > gpurun_out/r2_pcie_probe.txt 2>&1
python -m pytest tests/test_gpu_device_paths.py -m gpu -x -q -k "two_physical" 2>&1 | tail -3 > gpurun_out/r2_two_device_test.log
python tools/multi_gpu_solve.py --log2 23 > gpurun_out/r2_multi_gpu_solve.txt 2>&1
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -4 gpurun_out/r2_pcie_probe.txt | cut -c1-300; cat gpurun_out/r2_two_device_test.log; cat gpurun_out/r2_multi_gpu_solve.txt | grep -v "^\[" | cut -c1-200; head -c 400 gpurun_out/r2_bench_${N}gpu.json; tail -3 gpurun_out/r2_bench_${N}gpu.err
