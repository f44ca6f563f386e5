#!/bin/bash
# one 8-GPU box: topology, the host-link probe on all ranks at once (with and without binding), bench.py at N = 8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"; nvidia-smi topo -m) > gpurun_out/r2z_topo_n8.txt 2>&1
$TR --master-port 29511 scripts/pcie_probe.py --short > gpurun_out/r2z_pcie_n8.jsonl 2> gpurun_out/r2z_pcie_n8.err
$TR --master-port 29512 scripts/pcie_probe.py --short --bind > gpurun_out/r2z_pcie_n8_bind.jsonl 2> gpurun_out/r2z_pcie_n8_bind.err
$TR --master-port 29513 bench.py --gpus 8 > gpurun_out/r2z_bench_n8.json 2> gpurun_out/r2z_bench_n8.err
tail -c 600 gpurun_out/r2z_bench_n8.json
