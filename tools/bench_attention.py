"""A/B of the attention kernels on the model's shapes through the kernel-level C-ABI (CUDA events), each variant checked against a
materialised fp32 torch reference on one sequence: round-1 kernel (P through shared memory, 128-key tiles) vs the current one
(P in TMEM) at the per-call key tile width and at forced widths."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
from e2_tts_pytorch import _lib
from gpu_util import attention, DEV, rel

L = _lib.lib()
impl = C.c_int.in_dll(L, 'e2b_attention_impl')
force_bk = C.c_int.in_dll(L, 'e2b_attention_force_bk')
poly = C.c_int.in_dll(L, 'e2b_attention_poly')


def ref_one(q, k, v, kv_len, gate):
    sim = torch.einsum('hid,hjd->hij', q.float(), k.float())
    sim = torch.tanh(sim / 50.0) * 50.0
    sim[:, :, kv_len:] = -torch.finfo(torch.float32).max
    return torch.einsum('hij,hjd->hid', sim.softmax(-1), v.float()) * gate.permute(1, 0)[..., None]


def case(name, B, H, N, nkv, cross_mod=0, reps=5):
    g = torch.Generator().manual_seed(N + H)
    HD = H * 64
    kvb = cross_mod if cross_mod else B
    q = (torch.randn(B * N, HD, generator=g) * 0.3).to(torch.bfloat16).to(DEV)
    k = (torch.randn(kvb * nkv, HD, generator=g) * 1.5).to(torch.bfloat16).to(DEV)
    v = torch.randn(kvb, H, nkv, 64, generator=g).to(torch.bfloat16).to(DEV)
    npad = (nkv + 7) // 8 * 8
    vt = torch.zeros(kvb * H * 64, npad, device=DEV, dtype=torch.bfloat16)
    vt[:, :nkv] = v.permute(0, 1, 3, 2).reshape(kvb * H * 64, nkv)
    vrows = v.permute(0, 2, 1, 3).reshape(kvb * nkv, HD).contiguous()
    lens = torch.full((kvb,), nkv, device=DEV, dtype=torch.int32)
    lens[0] = max(1, nkv - 5)
    gate = torch.rand(B * N, H, generator=g).to(DEV)
    out = torch.zeros(B * N, HD, device=DEV, dtype=torch.bfloat16)
    kw = dict(batch=B, heads=H, q_rows_per_batch=N, kv_rows_per_batch=nkv, q=q, ldq=HD, q_col0=0, k=k, ldk=HD, k_col0=0, vt=vt, vt_ld=npad,
              kv_batch_mod=cross_mod, kv_lens=lens, kv_lens_add=0, hgate=gate, hgate_ld=H, out=out, ldo=HD, softclamp=50.0)
    r = ref_one(q[:N].reshape(N, H, 64).permute(1, 0, 2), k[:nkv].reshape(nkv, H, 64).permute(1, 0, 2), v[0], int(lens[0]), gate[:N])
    r = r.permute(1, 0, 2).reshape(N, HD)
    flops = 4.0 * B * H * N * nkv * 64
    row = f'{name:28s} B={B:4d} H={H:2d} N={N:5d} nkv={nkv:5d}:'
    # (label, kernel, forced key tile width, V as rows, quarter of the exponentials on the FMA pipe)
    variants = [('v1', 1, 0, 0, 0), ('v2 V^T', 0, 0, 0, 0), ('v2 rows', 0, 0, 1, 0), ('v2 V^T poly', 0, 0, 0, 1), ('v2 rows poly', 0, 0, 1, 1)] + \
        ([('v2 rows bk112', 0, 112, 1, 0)] if nkv == 782 else [])
    for label, im, bk, rows, pl in variants:
        impl.value, force_bk.value, poly.value = im, bk, pl
        kw.update(dict(vt=vrows, vt_ld=HD, v_rowmajor=1, v_col0=0) if rows else dict(vt=vt, vt_ld=npad, v_rowmajor=0, v_col0=0))
        out.zero_()
        attention(**kw)
        err = rel(out[:N], r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            attention(**kw)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        row += f'  {label} {us:8.1f} us ({flops / us / 1e6:5.0f} TF/s, err {err:.1e})'
    impl.value, force_bk.value, poly.value = 0, 0, 0
    print(row, flush=True)


case('self C2 audio/text 16h', 128, 16, 782, 782)
case('self C2 frames 8h', 128, 8, 782, 782)
case('self C4 30 s 16h', 32, 16, 2282, 2282)
case('self C5 20 s 16h', 64, 16, 1532, 1532)
case('self 5 s 16h', 128, 16, 407, 407)
case('cross C2 (T5, 8 keys)', 64, 16, 782, 8, cross_mod=64)
case('cross C4 (T5, 8 keys)', 16, 16, 2282, 8, cross_mod=16)
case('one clip 16h', 2, 16, 782, 782)
