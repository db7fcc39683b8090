"""Tuning driver: the fine NeRF MLP at full size, back to back for ~1.5 s per variant, with SM clock / power sampled by NVML
every 10 ms.  Run against the tuning build (NF_B200_LIB=.../libnf_b200_tune.so) to compare NF_MLP_IMPL / NF_MLP_CLUSTER."""
import os, sys, threading, time, torch, numpy as np
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import neurofluid_b200 as nb
from neurofluid_b200 import scenes, ops
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
net = nb.RenderNet(scenes.render_cfg(), 9., 13.); net.load_state_dict(scenes.init_render_state(0)); net = net.to(dev)
packed = ops.pack_nerf_weights(net.nerf_fine.ordered_params())
n = int(os.environ.get("ROWS", 11_200_000))
rec = torch.randn(n, 16, device=dev)

def run(label, env):
    for k in ("NF_MLP_IMPL", "NF_MLP_CLUSTER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(2):
        ops.nerf_mlp(packed, rec)
    torch.cuda.synchronize()
    samples, stop = [], False
    def sampler():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            time.sleep(0.01)
    th = threading.Thread(target=sampler); th.start()
    reps = int(os.environ.get("REPS", 100))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.nerf_mlp(packed, rec)
    e1.record(); torch.cuda.synchronize()
    stop = True; th.join()
    ms = e0.elapsed_time(e1) / reps
    s = samples[len(samples) // 4:]
    clk = np.array([x[0] for x in s]); pw = np.array([x[1] for x in s])
    reasons = 0
    for x in s:
        reasons |= x[2]
    print(f"{label}: {ms:.3f} ms/launch, {2 * 665984 * n / ms / 1e9:.0f} TFLOP/s | SM clock median {np.median(clk):.0f} min {clk.min()} max {clk.max()} MHz, "
          f"power median {np.median(pw):.0f} max {pw.max():.0f} W, throttle reasons 0x{reasons:x}, {len(s)} samples")

run("two tiles per CTA (production)", {})
if "tune" in os.path.basename(os.environ.get("NF_B200_LIB", "")):
    run("one tile per CTA", {"NF_MLP_IMPL": "1"})
    run("one tile per CTA, clusters of 4 with weight multicast", {"NF_MLP_IMPL": "1", "NF_MLP_CLUSTER": "4"})
    run("two tiles per CTA again", {})
