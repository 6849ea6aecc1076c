import time, torch, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(8192, 8192, device='cuda')
def busy():
    for _ in range(20): (x @ x)
for name, fn in [('clock', lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ('reasons', lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ('power', lambda: pynvml.nvmlDeviceGetPowerUsage(h)),
                 ('maxclock', lambda: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))]:
    ts = []
    for _ in range(5):
        busy()
        t = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t))
        torch.cuda.synchronize()
    print(name, ' '.join(f'{v:.2f}' for v in ts), 'ms')
