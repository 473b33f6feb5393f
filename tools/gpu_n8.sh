#!/bin/bash
# One 8-GPU session: host topology, PCIe contention between the ranks, bench.py as the driver launches it.
mkdir -p gpurun_out; O=gpurun_out
N=${1:-8}
(nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket"; free -g | head -2) > $O/r02_n${N}_topology.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
[ -z "$SKIP_PCIE" ] && timeout 200 $TR tools/pcie_contention.py > $O/r02_n${N}_pcie_contention.json 2> $O/r02_n${N}_pcie_contention.err
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 3 > $O/r02_scale_c4_n${N}.json 2> $O/r02_scale_c4_n${N}.err
tail -2 $O/r02_scale_c4_n${N}.err
python - <<PY
import json
d=json.load(open("$O/r02_scale_c4_n${N}.json"))
print("value",d["value"]/1e9,"e2e",d["e2e"]["value"]/1e9,"e2e_all",d["e2e_all_fields"]["value"]/1e9,"ms",d["ms_per_step"],d["config"]["state_checksum"],d["config"]["host_numa"])
c=json.load(open("$O/r02_n${N}_pcie_contention.json"))
print(c["aggregate"]); print(c["numa_nodes"]); print([r["gpu"] for r in c["ranks"]])
PY
