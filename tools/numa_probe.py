#!/usr/bin/env python
"""What the box tells about GPU <-> CPU / memory locality (sysfs and NVML), for bench.py's NUMA binding."""
import os
from pathlib import Path
import pynvml
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
nodes = sorted(p.name for p in Path("/sys/devices/system/node").glob("node[0-9]*"))
print("cpus allowed", len(os.sched_getaffinity(0)), "of", os.cpu_count(), "numa nodes", nodes)
for nd in nodes:
    print(nd, "cpulist", Path(f"/sys/devices/system/node/{nd}/cpulist").read_text().strip(),
          [l.strip() for l in Path(f"/sys/devices/system/node/{nd}/meminfo").read_text().splitlines()[:2]])
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    pci = pynvml.nvmlDeviceGetPciInfo(h)
    bus = pci.busId.decode() if isinstance(pci.busId, bytes) else pci.busId
    sysfs = Path(f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node")
    try:
        cpu = list(pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64))
    except Exception as ex:
        cpu = repr(ex)
    try:
        mem = list(pynvml.nvmlDeviceGetMemoryAffinity(h, 4, 0))
    except Exception as ex:
        mem = repr(ex)
    print("gpu", i, bus, "sysfs numa", sysfs.read_text().strip() if sysfs.exists() else None, "nvml cpu mask", [hex(x) for x in cpu] if isinstance(cpu, list) else cpu, "nvml mem nodes", mem)
