mkdir -p gpurun_out/topo
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; nproc; cat /sys/devices/system/node/online; free -g | head -2) > gpurun_out/topo/topo.txt 2>&1
python - >> gpurun_out/topo/topo.txt 2>&1 <<'PY'
import torch, os, glob
p = torch.cuda.get_device_properties(0)
print([a for a in dir(p) if 'pci' in a])
print(p.pci_bus_id, p.pci_device_id, p.pci_domain_id)
bus = "{:04x}:{:02x}:{:02x}.0".format(p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
print(bus)
for f in glob.glob(f"/sys/bus/pci/devices/{bus}/numa_node"): print(f, open(f).read())
print(os.sched_getaffinity(0))
PY
cat gpurun_out/topo/topo.txt
