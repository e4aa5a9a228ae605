"""Time one GEMM launch of the KTH step under the EXTDM_GEMM_DBG / EXTDM_HALO_ALL switches (profiling experiments)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import extdm_b200
from extdm_b200 import configs
key, B = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 32
model, cfg = configs.build_model("kth", device="cuda")
clip = torch.rand(B, 3, 10, 64, 64, device="cuda")
model.sample_one_video(1.0, clip)
runner = model.unet.runner(B, 32, 32, 16)
hit = None
for (fn, a, name), meta in zip(runner.step.steps, runner.step.meta):
    if name == "extdm_conv_gemm" and key in f"rows={meta['rows']} n={meta['n']} k={meta['k']} taps={meta['taps']}":
        hit = (fn, a, meta); break
fn, a, meta = hit
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3): fn(*a, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fn(*a, st)
e1.record(); torch.cuda.synchronize()
print(f"DBG={os.environ.get('EXTDM_GEMM_DBG','0')} HALO_ALL={'EXTDM_HALO_ALL' in os.environ} NO_HALO={'EXTDM_NO_HALO' in os.environ} B={B} rows={meta['rows']}: {e0.elapsed_time(e1)/10*1e3:.1f} us")
