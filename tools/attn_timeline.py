"""Per-tile timeline of CTA 0 of the attention kernel (clock64 stamps, see g_att_dbg in csrc/attention.cu)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
from e2_tts_pytorch import _lib
from gpu_util import attention, DEV
B, H, N = int(os.environ.get('BATCH', 16)), 16, 782
HD = H * 64
qk = (torch.randn(B * N, 2 * HD, device=DEV) * 0.5).to(torch.bfloat16)
VLD = int(os.environ.get('VLD', 784))
vt = torch.randn(B * H * 64, VLD, device=DEV).to(torch.bfloat16)
gate = torch.rand(B * N, H, device=DEV)
out = torch.empty(B * N, HD, device=DEV, dtype=torch.bfloat16)
lens = torch.full((B,), N, device=DEV, dtype=torch.int32)
kw = dict(batch=B, heads=H, q_rows_per_batch=N, kv_rows_per_batch=N, q=qk, ldq=2 * HD, q_col0=0, k=qk, ldk=2 * HD, k_col0=HD, vt=vt, vt_ld=VLD,
          kv_batch_mod=0, kv_lens=lens, kv_lens_add=0, hgate=gate, hgate_ld=H, out=out, ldo=HD, softclamp=50.0)
attention(**kw)
dbg = torch.zeros(96 * 8 + 8, device=DEV, dtype=torch.int64)
dbg[96 * 8] = int(os.environ.get('MODE', 1))
L = _lib.lib()
L.e2b_attention_set_debug.argtypes = [C.c_void_p]
assert L.e2b_attention_set_debug(C.c_void_p(dbg.data_ptr())) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); attention(**kw); e1.record(); torch.cuda.synchronize()
print(f'kernel {e0.elapsed_time(e1)*1e3:.1f} us for {B*H*7} items on 148 CTAs')
L.e2b_attention_set_debug(C.c_void_p(0))
t = dbg.cpu()[:96 * 8].reshape(96, 8)
t0 = int(t[0, 5])
names = ['sm_done', 'S_pre', 'PV_pre', 'PV_post', 'S_post', 'sm_arrive', 's_full_ok', 'p_empty_ok']
print('tile ' + ' '.join(f'{n:>10s}' for n in names) + ' | waitS waitP compute gap period')
prev = None
for g in range(int(os.environ.get('ROWS', 12))):
    row = [int(v) - t0 if int(v) else -1 for v in t[g]]
    done, arr, sok, pok = row[0], row[5], row[6], row[7]
    per = done - prev if prev is not None else 0
    gap = arr - prev if prev is not None else 0
    prev = done
    print(f'{g:4d} ' + ' '.join(f'{v:10d}' for v in row) + f' | {sok-arr:5d} {pok-sok:5d} {done-pok:7d} {gap:4d} {per:6d}')
