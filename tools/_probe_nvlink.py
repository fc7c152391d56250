import subprocess, sys
sys.path.insert(0, '.')
import bench
print("nvlink_kib(0):", bench.nvlink_kib(0))
try:
    import pynvml as nv
    nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
    for scope in (0, 0xFFFFFFFF):
        v = nv.nvmlDeviceGetFieldValues(h, [(nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, scope)])[0]
        print("scope", scope, "ret", v.nvmlReturn, "val", v.value.ullVal)
except Exception as e:
    print("nvml error", e)
print(subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], capture_output=True, text=True).stdout[:600])
