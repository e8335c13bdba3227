#!/bin/bash
# What the GPU box looks like from the host side (NUMA, PCIe, topology); outputs in gpurun_out/.
out=gpurun_out
mkdir -p $out
{
  echo "== nproc"; nproc
  echo "== lscpu"; lscpu | head -40
  echo "== numa"; ls /sys/devices/system/node/ 2>/dev/null; cat /sys/devices/system/node/node*/cpulist 2>/dev/null
  cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep -E "MemTotal|MemFree"
  which numactl && numactl -H
  echo "== free"; free -g
  echo "== topo"; nvidia-smi topo -m
  echo "== gpus"; nvidia-smi --query-gpu=index,name,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,clocks.max.sm,power.limit --format=csv
  for d in /sys/bus/pci/devices/*; do
    if [ -f $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi
  done
  echo "== cgroup"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
  echo "== pynvml"; python - <<'PY'
import time
import pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(0)
t0 = time.perf_counter()
for _ in range(100):
    c = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
    r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
print("sm", c, "reasons", hex(r), "per sample ms", (time.perf_counter() - t0) * 10)
PY
} > $out/box_probe.log 2>&1
