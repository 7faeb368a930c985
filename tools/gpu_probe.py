"""GPU bring-up probe: runs one named check per process (so a hung kernel can be killed by
`timeout` without taking the rest down) and prints PASS/FAIL lines.  Dev tool, not a test."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from transfusion_b200 import ops


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def gemm_case(M, N, K, a_mn, b_mn, tile_n=0, split_k=1, f32=False, bias=False, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    ref = A.float() @ B.float().t()
    bias_t = torch.randn(N, device=dev) if bias else None
    if bias:
        ref = ref + bias_t
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    if f32 or split_k > 1:
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a_st, b_st, out, M=M, N=N, K=K, a_mn_major=a_mn, b_mn_major=b_mn, tile_n=tile_n,
             split_k=split_k, accumulate=split_k > 1, bias=bias_t)
    torch.cuda.synchronize()
    e = rel(out.float(), ref)
    tag = f"gemm M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} tile_n={tile_n} split={split_k} f32={int(f32)}"
    print(("PASS " if e < 6e-3 else "FAIL ") + tag + f" rel={e:.3e}", flush=True)
    if e >= 6e-3:
        d = (out.float() - ref).abs()
        bad = (d > 0.05 * ref.abs().max()).nonzero()
        print("   first bad idx:", bad[:8].tolist(), " n_bad:", bad.shape[0], flush=True)
        # error map per 32x32 block
        Mb, Nb = min(M, 256), min(N, 256)
        blk = d[:Mb, :Nb].reshape(Mb // 32, 32, Nb // 32, 32).amax(dim=(1, 3)) if Mb % 32 == 0 and Nb % 32 == 0 else None
        if blk is not None:
            print("   blockmax(32x32):\n", (blk > 0.05 * ref.abs().max()).int().cpu().numpy(), flush=True)


def report(tag, e, tol):
    print(("PASS " if e < tol else "FAIL ") + tag + f" rel={e:.3e}", flush=True)


def attn_ref(q, k, v, H, d, kpm):
    B, Sq, _ = q.shape
    Sk = k.shape[1]
    qh = q.float().reshape(B, Sq, H, d).transpose(1, 2)
    kh = k.float().reshape(B, Sk, H, d).transpose(1, 2)
    vh = v.float().reshape(B, Sk, H, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / (d ** 0.5)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    lse2 = torch.logsumexp(s, dim=-1) * 1.4426950408889634
    o = torch.softmax(s, dim=-1) @ vh
    return o.transpose(1, 2).reshape(B, Sq, H * d), lse2


def attn_case(B, H, Sq, Sk, d, mask=False, fused=True, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    dp = (d + 31) // 32 * 32
    D = H * d
    if fused and Sq == Sk:
        qkv = torch.zeros(B * Sq, 3 * H * dp, device=dev, dtype=torch.bfloat16)
        src = torch.randn(B, Sq, 3, H, d, device=dev).bfloat16()
        qkv.view(B, Sq, 3, H, dp)[..., :d] = src
        q2, k2, v2 = qkv[:, :H * dp], qkv[:, H * dp:2 * H * dp], qkv[:, 2 * H * dp:]
        q, k, v = (src[:, :, i].reshape(B, Sq, D) for i in range(3))
    else:
        qs = torch.randn(B, Sq, H, d, device=dev).bfloat16()
        ks = torch.randn(B, Sk, H, d, device=dev).bfloat16()
        vs = torch.randn(B, Sk, H, d, device=dev).bfloat16()
        q2 = torch.zeros(B * Sq, H * dp, device=dev, dtype=torch.bfloat16); q2.view(B, Sq, H, dp)[..., :d] = qs
        k2 = torch.zeros(B * Sk, H * dp, device=dev, dtype=torch.bfloat16); k2.view(B, Sk, H, dp)[..., :d] = ks
        v2 = torch.zeros(B * Sk, H * dp, device=dev, dtype=torch.bfloat16); v2.view(B, Sk, H, dp)[..., :d] = vs
        q, k, v = qs.reshape(B, Sq, D), ks.reshape(B, Sk, D), vs.reshape(B, Sk, D)
    kpm = None
    if mask:
        kpm = torch.zeros(B, Sk, dtype=torch.uint8, device=dev)
        for b in range(B):
            kpm[b, Sk - 1 - 7 * b - 3:] = 1
    out = torch.zeros(B * Sq, H * dp, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, Sq, device=dev)
    ops.attn_fwd(q2, k2, v2, out, lse, B=B, H=H, Sq=Sq, Sk=Sk, dp=dp, scale=1.0 / d ** 0.5,
                 key_padding_mask=kpm, kpm_start=max(0, Sk - 64))
    torch.cuda.synchronize()
    ref, lse_ref = attn_ref(q, k, v, H, d, kpm)
    got = out.view(B, Sq, H, dp)[..., :d].reshape(B, Sq, D).float()
    tag = f"attn_fwd B={B} H={H} Sq={Sq} Sk={Sk} d={d} mask={int(mask)} fused={int(fused)}"
    report(tag + " out", rel(got, ref), 1e-2)
    report(tag + " lse", float((lse - lse_ref).abs().max() / lse_ref.abs().max()), 1e-3)
    if dp != d:
        padmax = float(out.view(B, Sq, H, dp)[..., d:].abs().max())
        print("   pad-cols max:", padmax, flush=True)


def attn_bwd_case(B, H, S, d, mask=True, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    dp = (d + 31) // 32 * 32
    D = H * d
    src = torch.randn(B, S, 3, H, d, device=dev).bfloat16()
    qkv = torch.zeros(B * S, 3 * H * dp, device=dev, dtype=torch.bfloat16)
    qkv.view(B, S, 3, H, dp)[..., :d] = src
    q2, k2, v2 = qkv[:, :H * dp], qkv[:, H * dp:2 * H * dp], qkv[:, 2 * H * dp:]
    kpm = None
    if mask:
        kpm = torch.zeros(B, S, dtype=torch.uint8, device=dev)
        for b in range(B):
            kpm[b, S - 1 - 5 * b - 2:] = 1
    Sp = (S + 127) // 128 * 128
    out = torch.zeros(B * S, H * dp, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, Sp, device=dev)
    ops.attn_fwd(q2, k2, v2, out, lse, B=B, H=H, Sq=S, Sk=S, dp=dp, scale=1.0 / d ** 0.5, key_padding_mask=kpm, kpm_start=0)
    dout_src = torch.randn(B, S, H, d, device=dev).bfloat16()
    dout = torch.zeros(B * S, H * dp, device=dev, dtype=torch.bfloat16)
    dout.view(B, S, H, dp)[..., :d] = dout_src
    delta = torch.zeros(B, H, Sp, device=dev)
    ops.attn_delta(out, dout, delta, B, S, H, dp)
    dqkv = torch.full((B * S, 3 * H * dp), 7.0, device=dev, dtype=torch.bfloat16)
    ops.attn_bwd(q2, k2, v2, dout, lse, delta, dqkv[:, :H * dp], dqkv[:, H * dp:2 * H * dp], dqkv[:, 2 * H * dp:],
                 B=B, H=H, Sq=S, Sk=S, dp=dp, scale=1.0 / d ** 0.5, key_padding_mask=kpm)
    torch.cuda.synchronize()
    # reference
    qf, kf, vf = (src[:, :, i].float().reshape(B, S, D).requires_grad_(True) for i in range(3))
    ref, _ = attn_ref(qf, kf, vf, H, d, kpm)
    ref.backward(dout_src.float().reshape(B, S, D))
    got = dqkv.view(B, S, 3, H, dp)
    tag = f"attn_bwd B={B} H={H} S={S} d={d} mask={int(mask)}"
    dref = (out.view(B, S, H, dp)[..., :d].float() * dout_src.float()).sum(-1).permute(0, 2, 1)
    report(tag + " delta", rel(delta[:, :, :S], dref), 1e-4)
    for i, (nm, g) in enumerate((("dq", qf.grad), ("dk", kf.grad), ("dv", vf.grad))):
        report(tag + " " + nm, rel(got[:, :, i, :, :d].reshape(B, S, D).float(), g), 1.5e-2)
    if dp != d:
        print("   pad max:", float(got[..., d:].float().abs().max()), flush=True)


def ln_case(rows, D, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    x = torch.randn(rows, D, device=dev).bfloat16()
    g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
    y = torch.empty_like(x); mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
    ops.layernorm_fwd(x, y, g, b, mean, rstd, rows, D)
    xr = x.float().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), g, b, 1e-5)
    report(f"ln_fwd rows={rows} D={D}", rel(y.float(), yr), 6e-3)
    dy = torch.randn(rows, D, device=dev).bfloat16()
    gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    yr2 = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
    yr2.backward(dy.float())
    dx = torch.empty_like(x); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); dbias = torch.zeros(D, device=dev)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx, dg, db, rows, D, dbias=dbias)
    torch.cuda.synchronize()
    report(f"ln_bwd dx rows={rows} D={D}", rel(dx.float(), xr.grad), 8e-3)
    report(f"ln_bwd dgamma", rel(dg, gr.grad), 1e-3)
    report(f"ln_bwd dbeta", rel(db, br.grad), 1e-3)
    report(f"ln_bwd dbias", rel(dbias, dx.float().sum(0)), 1e-3)
    # remapped: read only first n of every S rows, write compact
    B, S, n = 3, rows // 3, rows // 3 - 5
    yc = torch.empty(B * n, D, device=dev, dtype=torch.bfloat16)
    ops.layernorm_fwd(x, yc, g, b, None, None, B * n, D, in_map=(n, S, 0))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float().view(B, S, D)[:, :n], (D,), g, b, 1e-5).reshape(B * n, D)
    report(f"ln_fwd remap", rel(yc.float(), ref), 6e-3)


def layout_case(B, Cc, H, W, p, dtype=torch.float32):
    import sys as _s
    torch.manual_seed(0)
    dev = "cuda"
    f = torch.randn(B, Cc, H, W, device=dev).to(dtype)
    gh, gw = H // p, W // p
    tok = torch.empty(B * gh * gw, Cc * p * p, device=dev, dtype=torch.bfloat16)
    ops.patchify(f, p, tok)
    ref = f.float().reshape(B, Cc, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, Cc * p * p).bfloat16()
    torch.cuda.synchronize()
    ok = torch.equal(tok, ref)
    print(("PASS " if ok else "FAIL ") + f"patchify B={B} C={Cc} {H}x{W} p={p} {dtype}", flush=True)
    out = torch.zeros(B, Cc, H, W, device=dev, dtype=dtype)
    ops.fold(tok, out, p)
    torch.cuda.synchronize()
    ok = torch.equal(out.float(), f.bfloat16().float())
    print(("PASS " if ok else "FAIL ") + f"fold B={B} C={Cc} {H}x{W} p={p} {dtype}", flush=True)


def misc_case():
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(1000, 2688, device=dev).bfloat16()
    out = torch.zeros(2688, device=dev)
    ops.colsum(x, out, 1000, 2688)
    report("colsum", rel(out, x.float().sum(0)), 1e-4)
    B, L, D, n = 3, 7, 64, 10
    S = n + L
    lang = torch.randn(B, L, D, device=dev); kind = torch.randn(D, device=dev)
    z = torch.zeros(B, S, D, device=dev, dtype=torch.bfloat16)
    ops.lang_rows_fwd(lang, kind, z, n)
    report("lang_rows_fwd", rel(z[:, n:].float(), (lang + kind).bfloat16().float()), 1e-6)
    dz = torch.randn(B, S, D, device=dev).bfloat16()
    dlang = torch.zeros(B, L, D, device=dev); dkind = torch.zeros(D, device=dev)
    ops.lang_rows_bwd(dz, dlang, dkind, B, L, n)
    report("lang_rows_bwd dlang", rel(dlang, dz[:, n:].float()), 1e-6)
    report("lang_rows_bwd dkind", rel(dkind, dz[:, n:].float().sum((0, 1))), 1e-5)
    # cast_pad: rows 3 blocks of 10 -> 16
    w = torch.randn(30, 24, device=dev)
    wp = torch.zeros(48, 24, device=dev, dtype=torch.bfloat16)
    ops.cast_pad(w, wp, 30, 24, rin=10, rout=16)
    ok = torch.equal(wp.view(3, 16, 24)[:, :10].reshape(30, 24), w.bfloat16()) and float(wp.view(3, 16, 24)[:, 10:].abs().max()) == 0
    print(("PASS " if ok else "FAIL ") + "cast_pad rows", flush=True)
    wp2 = torch.zeros(30, 3 * 16, device=dev, dtype=torch.bfloat16)
    w2 = torch.randn(30, 30, device=dev)
    ops.cast_pad(w2, wp2, 30, 30, cin=10, cout=16)
    ok = torch.equal(wp2.view(30, 3, 16)[:, :, :10].reshape(30, 30), w2.bfloat16())
    print(("PASS " if ok else "FAIL ") + "cast_pad cols", flush=True)
    gsrc = torch.randn(30, 48, device=dev); gdst = torch.ones(30, 30, device=dev)
    ops.unpad_add(gsrc, gdst, 30, 30, cin=10, cout=16)
    ok = torch.allclose(gdst, 1 + gsrc.view(30, 3, 16)[:, :, :10].reshape(30, 30))
    print(("PASS " if ok else "FAIL ") + "unpad_add cols", flush=True)
    o = torch.randn(50, 4 * 32, device=dev).bfloat16(); do = torch.randn(50, 4 * 32, device=dev).bfloat16()
    delta = torch.zeros(50, 4, device=dev)
    ops.attn_delta(o, do, delta, 50, 4, 32)
    report("attn_delta", rel(delta, (o.float() * do.float()).view(50, 4, 32).sum(-1)), 1e-5)


def level_debug(name="c4_d40_oddhead", level=0):
    """Stage-by-stage comparison of one level of a golden case against a torch fp32 restatement."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from tests.golden_utils import load_golden
    from tests.fusion_testlib import build_module
    from transfusion_b200.cross_fusion import level_fn
    from oracle import ref_math
    g = load_golden(name)
    feats = g["features"]
    keys = sorted(feats, key=int)
    shapes = [tuple(feats[k].shape[2:]) for k in keys]
    channels = [feats[k].shape[1] for k in keys]
    D = g["lang"].shape[-1]
    m = build_module(D, shapes, channels, g["patch"], g["layers"], g["heads"], lm=g["lm_on"])
    m.load_state_dict(g["params"], strict=False)
    m.train()
    sink = {}
    level_fn.DEBUG_SINK = sink
    i = level
    feat = feats[keys[i]].cuda()
    lang = g["lang"].cuda()
    pad = ~(g["att_mask"].bool()).cuda()
    with torch.no_grad():
        fused, _ = m.run_level(i, feat, lang, pad)
    torch.cuda.synchronize()
    level_fn.DEBUG_SINK = None
    # reference intermediates (fp32, on GPU)
    sd = {k: v.cuda() for k, v in g["params"].items()}
    p = g["patch"][i]; H = g["heads"]; nl = g["layers"][i]
    B, C, h, w = feat.shape
    n = (h // p) * (w // p)
    enc = f"cross_fusion_encoders.{i}."
    tok = ref_math.patchify(feat, p)
    report("tok", rel(sink["tok"].view_as(tok), tok), 5e-3)
    x = tok @ sd[f"patches_to_token.{i}.weight"].reshape(D, -1).t() + ref_math.sin1d_table(n, D).cuda() + sd[enc + "image_kind_embedding"].reshape(D)
    lg = lang + sd[enc + "lang_kind_embedding"].reshape(D)
    z = torch.cat([x, lg], 1)
    report("z0", rel(sink["z0"], z), 5e-3)
    key_pad = torch.cat([torch.zeros(B, n, dtype=torch.bool, device="cuda"), pad], 1)
    d = D // H; dp = (d + 31) // 32 * 32
    S = z.shape[1]
    for l in range(nl):
        pre = enc + f"t_encoder.layers.{l}."
        qkv = z @ sd[pre + "self_attn.in_proj_weight"].t() + sd[pre + "self_attn.in_proj_bias"]
        got = sink[f"l{l}.qkv"].view(B, S, 3, H, dp)
        report(f"l{l}.qkv", rel(got[..., :d].reshape(B, S, 3 * D), qkv), 5e-3)
        print("    qkv pad max", float(got[..., d:].abs().max()) if dp != d else 0.0)
        q, k, v = qkv.split(D, -1)
        o = ref_math.attention(q, k, v, key_pad, H)
        got = sink[f"l{l}.att"].view(B, S, H, dp)[..., :d].reshape(B, S, D)
        report(f"l{l}.att", rel(got, o), 8e-3)
        y1 = z + o @ sd[pre + "self_attn.out_proj.weight"].t() + sd[pre + "self_attn.out_proj.bias"]
        report(f"l{l}.y1", rel(sink[f"l{l}.y1"].view_as(y1), y1), 8e-3)
        x1 = ref_math.layer_norm(y1, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
        report(f"l{l}.x1", rel(sink[f"l{l}.x1"].view_as(x1), x1), 8e-3)
        hh = ref_math.gelu_erf(x1 @ sd[pre + "linear1.weight"].t() + sd[pre + "linear1.bias"])
        report(f"l{l}.h", rel(sink[f"l{l}.h"].view_as(hh), hh), 8e-3)
        y2 = x1 + hh @ sd[pre + "linear2.weight"].t() + sd[pre + "linear2.bias"]
        report(f"l{l}.y2", rel(sink[f"l{l}.y2"].view_as(y2), y2), 8e-3)
        z = ref_math.layer_norm(y2, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
        report(f"l{l}.x2", rel(sink[f"l{l}.x2"].view_as(z), z), 8e-3)
    vis = ref_math.layer_norm(z[:, :n], sd[enc + "final_norm_layer.weight"], sd[enc + "final_norm_layer.bias"])
    report("vis", rel(sink["vis"].view_as(vis), vis), 8e-3)
    yb = vis @ sd[f"tokens_to_features.{i}.linear.weight"].t() + sd[f"tokens_to_features.{i}.linear.bias"]
    report("yb", rel(sink["yb"].view_as(yb), yb), 8e-3)
    out = ref_math.fold(yb, C, p, h // p, w // p)
    report("fused", rel(fused.float(), out), 8e-3)
    report("fused vs golden", rel(fused.float().cpu(), g["out"][keys[i]]), 8e-3)


CASES = {
    "dbg_c4": lambda: level_debug("c4_d40_oddhead", 0),
    "dbg_f4_l0": lambda: level_debug("fusion4_d32", 0),
    "dbg_f4_l3": lambda: level_debug("fusion4_d32", 3),
    "dbg_c5": lambda: level_debug("c5_d64_lm", 0),
    "attn_small": lambda: attn_case(1, 1, 128, 64, 64),
    "attn_2tiles": lambda: attn_case(1, 1, 128, 128, 64),
    "attn_multi": lambda: attn_case(2, 4, 300, 300, 64, mask=True),
    "attn_d224": lambda: attn_case(2, 4, 832, 832, 224, mask=True),
    "attn_d178": lambda: attn_case(2, 4, 500, 500, 178, mask=True),
    "attn_cross": lambda: attn_case(2, 2, 200, 333, 32, mask=True, fused=False),
    "attn_d16": lambda: attn_case(2, 4, 70, 70, 16, mask=True),
    "attn_big": lambda: attn_case(2, 4, 3136, 3136, 224, mask=True),
    "attn_bwd_small": lambda: attn_bwd_case(1, 1, 128, 32, mask=False),
    "attn_bwd_s64": lambda: attn_bwd_case(1, 1, 64, 64, mask=False),
    "attn_bwd_multi": lambda: attn_bwd_case(2, 2, 300, 64),
    "attn_bwd_d224": lambda: attn_bwd_case(2, 4, 832, 224),
    "attn_bwd_d178": lambda: attn_bwd_case(2, 4, 500, 178),
    "attn_bwd_d16": lambda: attn_bwd_case(2, 4, 70, 16),
    "attn_bwd_big": lambda: attn_bwd_case(1, 4, 3136, 224),
    "ln": lambda: (ln_case(999, 896), ln_case(300, 712), ln_case(129, 64)),
    "layout": lambda: (layout_case(2, 256, 16, 24, 4), layout_case(2, 512, 8, 12, 4, torch.bfloat16),
                       layout_case(2, 64, 12, 20, 2), layout_case(3, 2048, 24, 32, 1), layout_case(2, 8, 16, 24, 4)),
    "misc": misc_case,
    "gemm_kk_small": lambda: gemm_case(128, 128, 64, False, False, tile_n=128),
    "gemm_kk_k256": lambda: gemm_case(128, 128, 256, False, False, tile_n=128),
    "gemm_kk_multi": lambda: gemm_case(1024, 896, 896, False, False, bias=True),
    "gemm_kk_tail": lambda: gemm_case(1000, 712, 712, False, False, bias=True),
    "gemm_kk_big": lambda: gemm_case(8192, 2688, 896, False, False),
    "gemm_kmn_small": lambda: gemm_case(128, 128, 64, False, True, tile_n=128),
    "gemm_kmn": lambda: gemm_case(512, 896, 1792, False, True),
    "gemm_mnk_small": lambda: gemm_case(128, 128, 64, True, False, tile_n=128),
    "gemm_mnmn_small": lambda: gemm_case(128, 128, 128, True, True, tile_n=128),
    "gemm_mnmn": lambda: gemm_case(896, 1792, 4096, True, True, split_k=4),
    "gemm_mnmn_tail": lambda: gemm_case(712, 1424, 1000, True, True, split_k=3),
    "gemm_f32": lambda: gemm_case(256, 256, 512, False, False, f32=True),
}



def drop_debug():
    torch.manual_seed(3)
    M, N, K, p = 2048, 896, 64, 0.1
    A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
    outs = []
    for first in (False, True, False, True):
        o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(A, W, o, M=M, N=N, K=K, drop_p=p, drop_seed=123, drop_stream=7, drop_first=first)
        torch.cuda.synchronize()
        outs.append(o != 0)
    for i in range(1, 4):
        diff = outs[0] != outs[i]
        idx = diff.nonzero()
        print(f"run{i} vs run0: mismatches {int(diff.sum())}", idx[:10].tolist(), flush=True)
        if idx.numel():
            print("   rows mod 128:", sorted(set((idx[:, 0] % 128).tolist()))[:20], " cols mod 32:", sorted(set((idx[:, 1] % 32).tolist()))[:32])


def drop_debug2():
    torch.manual_seed(3)
    M, N, K, p = 2048, 896, 64, 0.1
    A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
    ref = A.float() @ W.float().t()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, out, M=M, N=N, K=K, drop_p=p, drop_seed=123, drop_stream=7)
    keep = out != 0
    A2 = (torch.rand(M, K, device="cuda") + 0.5).bfloat16(); W2 = (torch.rand(N, K, device="cuda") + 0.5).bfloat16()
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A2, W2, out2, M=M, N=N, K=K, drop_p=p, drop_seed=123, drop_stream=7, drop_first=True)
    torch.cuda.synchronize()
    k2 = out2 != 0
    diff = keep != k2
    idx = diff.nonzero()
    print("mismatches", int(diff.sum()), idx[:10].tolist())
    for (r, c) in idx[:10].tolist():
        print("  at", r, c, "out", float(out[r, c]), "ref", float(ref[r, c]), "out2", float(out2[r, c]))


def t0b(mma):
    return int(mma[10, 0])


def bwd_timeline(drop=0.15):
    import math
    B, H, S, d = 13, 4, 3136, 224
    dev = "cuda"
    torch.manual_seed(0)
    qkv = torch.randn(B * S, 3 * H * d, device=dev).bfloat16()
    D = H * d
    Sp = (S + 127) // 128 * 128
    out = torch.empty(B * S, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Sp, device=dev)
    kpm = torch.zeros(B, S, dtype=torch.uint8, device=dev); kpm[:, S - 20:] = 1
    kw = dict(B=B, H=H, Sq=S, Sk=S, dp=d, scale=1 / math.sqrt(d), drop_p=drop, drop_seed=3, drop_stream=4)
    ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, lse, key_padding_mask=kpm, kpm_start=S - 64, **kw)
    dout = torch.randn(B * S, D, device=dev).bfloat16()
    delta = torch.empty(B, H, Sp, device=dev)
    ops.attn_delta(out, dout, delta, B, S, H, d)
    dqkv = torch.empty(B * S, 3 * D, device=dev, dtype=torch.bfloat16)
    dbg = torch.zeros(3, 2, 64, 8, dtype=torch.int64, device=dev)
    for it in range(2):
        ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], dout, lse, delta, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                     key_padding_mask=kpm, kpm_start=S - 64, debug_timeline=dbg, **kw)
    torch.cuda.synchronize()
    t = dbg.cpu()
    for pi, pname in enumerate(("dQ pass", "dK pass", "dV pass")):
        mma, sm = t[pi, 0], t[pi, 1]
        t0 = int(mma[8, 0])
        print(f"--- {pname} drop={drop}: iteration period (MMA thread) cycles:", [int(mma[i + 1, 0] - mma[i, 0]) for i in range(8, 24)])
        print("MMA thread per-iter deltas [L_FULL wait, S_FULL wait, C_EMPTY wait, issue C, (E_FULL wait), issue acc] (iters 10..17):")
        for i in range(10, 18):
            r = mma[i]
            print("   ", int(r[6] - r[0]), int(r[1] - r[6]), int(r[2] - r[1]), int(r[3] - r[2]), int(r[4] - r[3]), int(r[5] - r[4]))
        print("element-wise warp0 per-iter deltas [C_FULL wait, ld+compute, E_EMPTY wait, st+arrive]:")
        for i in range(10, 18):
            r = sm[i]
            print("   ", int(r[1] - r[0]), int(r[3] - r[1]), int(r[4] - r[3]), int(r[5] - r[4]), " period", int(sm[i + 1, 0] - r[0]))
        # absolute timeline (same SM clock) of iterations 10..12, relative to the MMA thread's start of iteration 10
        ev = []
        names_m = {0: "M loop top", 6: "M long-ring tile full", 1: "M tiles full", 2: "M C_EMPTY passed -> issue C", 3: "M C issued+committed", 4: "M E_FULL(i-1) passed -> issue acc(i-1)", 5: "M acc issued"}
        names_e = {0: "E loop top (stats requested)", 1: "E C_FULL passed", 2: "E w0 C_EMPTY arrived", 6: "E w7 C_EMPTY arrived", 3: "E compute done", 4: "E E_EMPTY passed", 5: "E E stored+arrived"}
        for i in range(10, 13):
            for k, nm in names_m.items():
                ev.append((int(mma[i, k]) - t0b(mma), f"i={i} {nm}"))
            for k, nm in names_e.items():
                ev.append((int(sm[i, k]) - t0b(mma), f"i={i} {nm}"))
        for tt, nm in sorted(ev):
            print(f"      {tt:7d}  {nm}")


def fwd_timeline(drop=0.15):
    import math
    B, H, S, d = 13, 4, 3136, 224
    dev = "cuda"
    torch.manual_seed(0)
    qkv = torch.randn(B * S, 3 * H * d, device=dev).bfloat16()
    D = H * d
    Sp = (S + 127) // 128 * 128
    out = torch.empty(B * S, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Sp, device=dev)
    kpm = torch.zeros(B, S, dtype=torch.uint8, device=dev); kpm[:, S - 20:] = 1
    dbg = torch.zeros(2, 64, 8, dtype=torch.int64, device=dev)
    kw = dict(B=B, H=H, Sq=S, Sk=S, dp=d, scale=1 / math.sqrt(d), drop_p=drop, drop_seed=3, drop_stream=4)
    for it in range(2):
        ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, lse, key_padding_mask=kpm, kpm_start=S - 64,
                     debug_timeline=dbg, **kw)
    torch.cuda.synchronize()
    t = dbg.cpu()
    mma, sm = t[0], t[1]
    print(f"--- attn fwd drop={drop}: MMA-thread period per KV tile:", [int(mma[i + 1, 0] - mma[i, 0]) for i in range(8, 20)])
    print("MMA thread [K_FULL wait, S_EMPTY wait, issue S | (gap) P_FULL wait, V_FULL wait, issue PV]:")
    for i in range(10, 16):
        r = mma[i]
        print("   ", int(r[1] - r[0]), int(r[2] - r[1]), int(r[3] - r[2]), "|", int(r[5] - r[4]), int(r[6] - r[5]), int(r[7] - r[6]))
    print("softmax warp2 [S_FULL wait, ld+release, compute, O_READY wait, rescale+P store+arrive]:")
    for i in range(10, 16):
        r = sm[i]
        print("   ", int(r[1] - r[0]), int(r[2] - r[1]), int(r[3] - r[2]), int(r[4] - r[3]), int(r[5] - r[4]), " period", int(sm[i + 1, 0] - r[0]))


CASES["fwd_timeline"] = lambda: (fwd_timeline(0.15), fwd_timeline(0.0))
CASES["bwd_timeline"] = lambda: (bwd_timeline(0.15), bwd_timeline(0.0))
CASES["drop_debug2"] = drop_debug2
CASES["drop_debug"] = drop_debug


if __name__ == "__main__":
    name = sys.argv[1]
    if name == "list":
        print(" ".join(CASES))
    else:
        CASES[name]()
