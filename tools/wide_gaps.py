"""Wall-clock (globaltimer) gaps between back-to-back wide forward launches in a CUDA graph."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SRST_ST_DEBUG"] = "1"
os.environ.setdefault("SRST_ST_FWD_CFG", "10")
from srgan_st_b200 import _cabi, taps as T
lib = _cabi.lib(); g, dg = T.gaussian_taps(0.5); k, _ = T.gaussian_taps(2.0)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
B, H, W = 64, 96, 96
N = 6
srs = [torch.rand(B, 3, H, W, device="cuda") for _ in range(N)]; hrs = [torch.rand(B, 3, H, W, device="cuda") for _ in range(N)]
ds = torch.empty_like(srs[0]); loss = torch.zeros((), device="cuda")
wss = [torch.zeros(65536, dtype=torch.uint8, device="cuda") for _ in range(N)]
st = torch.cuda.Stream(); sp = ctypes.c_void_p(st.cuda_stream)
def fwd(i, s):
    _cabi.check(lib.srst_st_forward(vp(srs[i]), vp(hrs[i]), B, H, W, T.as_c(g), T.as_c(dg), 2, T.as_c(k), 8, 1, 1e-12, vp(loss), vp(ds), None, None, None, vp(wss[i]), wss[i].numel(), s), "fwd")
fwd(0, None); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr, stream=st):
    for i in range(N): fwd(i, sp)
for _ in range(3):
    gr.replay(); torch.cuda.synchronize()
rows = []
for i in range(N):
    d = wss[i][4096:4096 + 8 * 64].view(torch.int64).cpu()
    rows.append([int(d[24]), int(d[25]), int(d[26]), int(d[27]), int(d[32 + 24]), int(d[32 + 25]), int(d[32 + 26]), int(d[32 + 27])])
t0 = rows[0][0]
print("launch: cta0[entry, after-wait, stores-done, exit]  lastcta[entry, after-wait, stores-done, exit]  (ns from first entry)")
for i, r in enumerate(rows):
    print(i, [x - t0 if x else None for x in r])
