set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c4
nvidia-smi topo -m > gpurun_out/c4/topo.txt 2>&1
timeout 1200 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/c4/pytest_multigpu.txt 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/c4/pytest_multigpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/bandwidth/sweep.py --out gpurun_out/c4/bandwidth_sweep_n2.json > gpurun_out/c4/sweep_n2.txt 2>&1; echo "sweep rc=$?"
tail -15 gpurun_out/c4/sweep_n2.txt
