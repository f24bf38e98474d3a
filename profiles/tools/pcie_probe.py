"""H2D copy-engine bandwidth from pinned memory (CUDA events) for the sizes the host-buffer step moves,
and whether stream memory operations (cuStreamWriteValue32) are usable on this box."""
import torch

torch.cuda.init()
dev = torch.device("cuda", 0)
for size in (64 << 10, 256 << 10, 512 << 10, 2 << 20, 16 << 20):
    h = torch.empty(size, dtype=torch.uint8).pin_memory()
    d = torch.empty(size, dtype=torch.uint8, device=dev)
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"H2D {size >> 10:6d} KiB: {us:8.1f} us  {size / us / 1e3:6.1f} GB/s", flush=True)
try:
    from cuda.bindings import driver as cu
except Exception:  # older cuda-python layout
    from cuda import cuda as cu
err, = cu.cuInit(0)
err, cdev = cu.cuDeviceGet(0)
for name in ("CU_DEVICE_ATTRIBUTE_CAN_USE_STREAM_WAIT_VALUE_NOR", "CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS",
             "CU_DEVICE_ATTRIBUTE_CAN_USE_STREAM_MEM_OPS_V1", "CU_DEVICE_ATTRIBUTE_CAN_FLUSH_REMOTE_WRITES",
             "CU_DEVICE_ATTRIBUTE_PCI_BUS_ID"):
    a = getattr(cu.CUdevice_attribute, name, None)
    if a is not None:
        print(name, cu.cuDeviceGetAttribute(a, cdev), flush=True)
flag = torch.zeros(4, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
r = cu.cuStreamWriteValue32(s, flag.data_ptr(), 7, 0)
torch.cuda.synchronize()
print("cuStreamWriteValue32 ->", r, "flag =", flag.tolist(), flush=True)
import subprocess
print(subprocess.run("nvidia-smi -q | grep -A12 'GPU Link Info' | head -16", shell=True, capture_output=True, text=True).stdout)
