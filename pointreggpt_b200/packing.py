"""Weight packing: reference state-dict -> the blob libprg.so consumes.

Done once at load time on the host (weights are frozen during generation):
  * weight standardisation (SDD:606-613, fp32 branch, eps 1e-5) is folded into the stored
    conv matrices -- the reference recomputes it in every forward of its 38 WS convs;
  * conv weights (Cout, Cin, kh, kw) become fp16 K-major "tap-major" matrices
    (Cout, kh*kw*Cin) -- the B operand of the implicit GEMM;
  * nearest-x2 Upsample + conv3x3 (SDD:592-594) becomes four 2x2-tap parity classes with
    pre-summed taps (2.25x fewer MACs, no materialised upsampled tensor);
  * everything else (biases, GroupNorm affine, LayerNorm gains, MLPs, stem, final 1x1) stays fp32.

Blob format (little endian): magic 'PRGW', u32 version, u32 n_entries, then per entry
  u16 name_len, name bytes, u8 dtype (0=f32, 1=f16), u8 ndim, u32 dims[ndim], u64 offset,
  u64 nbytes; payload 256-byte aligned after the table.
"""
import struct

import numpy as np
import torch

MAGIC = b"PRGW"
VERSION = 1


def standardize(w, eps=1e-5):
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (w - mean) * (var + eps).rsqrt()


def conv_weight_kmajor(w):
    """(Cout, Cin, kh, kw) -> (Cout, kh*kw*Cin), k = (ky*kw + kx)*Cin + c."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous()


def upsample_fold_weight(w):
    """3x3 conv applied after nearest-x2 upsampling == per output parity (py, px) a 2x2 conv
    on the low-res input.  Returns (Cout, 4 classes * 4 taps * Cin); class = py*2+px,
    tap = ty*2+tx, source row offset = ty - (1 - py) (same for columns)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    sel = {0: ([0], [1, 2]), 1: ([0, 1], [2])}   # parity -> original taps feeding tap 0 / tap 1
    out = torch.zeros(co, 4, 4, ci, dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = torch.zeros(co, ci, dtype=torch.float32)
                    for ky in sel[py][ty]:
                        for kx in sel[px][tx]:
                            acc += w[:, :, ky, kx].float()
                    out[:, py * 2 + px, ty * 2 + tx] = acc
    return out.reshape(co, 16 * ci).contiguous()


def upsample_rows3_weight(w):
    """The folded upsample conv once more, for the class-bound row-streaming kernel (conv2_tc.cu, rows3):
    per parity class a 3x3 conv on the input shifted by (py, px) whose ky = 2 / kx = 2 taps are zero.
    Returns (Cout, 4 classes * 9 taps * Cin), k = ((cls * 3 + ky) * 3 + kx) * Cin + c."""
    co, ci = w.shape[0], w.shape[1]
    f = upsample_fold_weight(w).reshape(co, 4, 2, 2, ci)
    out = torch.zeros(co, 4, 3, 3, ci, dtype=torch.float32)
    out[:, :, :2, :2] = f
    return out.reshape(co, 36 * ci).contiguous()


class BlobWriter:
    def __init__(self):
        self.entries = []

    def add(self, name, tensor, dtype):
        t = tensor.detach().to(torch.float32).cpu().contiguous()
        if dtype == "f16":
            arr = t.to(torch.float16).numpy()
            code = 1
        else:
            arr = t.numpy().astype(np.float32)
            code = 0
        self.entries.append((name, code, arr))

    def tobytes(self):
        table = []
        size = 12
        for name, code, arr in self.entries:
            size += 2 + len(name.encode()) + 2 + 4 * arr.ndim + 16
        off = (size + 255) // 256 * 256
        payload = []
        for name, code, arr in self.entries:
            nb = arr.nbytes
            table.append((name.encode(), code, arr.shape, off, nb))
            payload.append((off, arr.tobytes()))
            off = (off + nb + 255) // 256 * 256
        buf = bytearray(off)
        struct.pack_into("<4sII", buf, 0, MAGIC, VERSION, len(table))
        p = 12
        for name, code, shape, o, nb in table:
            struct.pack_into("<H", buf, p, len(name)); p += 2
            buf[p:p + len(name)] = name; p += len(name)
            struct.pack_into("<BB", buf, p, code, len(shape)); p += 2
            for d in shape:
                struct.pack_into("<I", buf, p, d); p += 4
            struct.pack_into("<QQ", buf, p, o, nb); p += 16
        for o, b in payload:
            buf[o:o + len(b)] = b
        return bytes(buf)


# --------------------------------------------------------------------------- #
# state-dict -> blob
# --------------------------------------------------------------------------- #
KIND_UNET, KIND_MASKUNET = 1, 2


def _levels(sd):
    n = 0
    while ("downs.%d.0.block1.proj.weight" % n) in sd:
        n += 1
    return n


def resblock_order(levels):
    """Order in which the forward pass visits the ResnetBlocks (= row order of mlp_all)."""
    names = []
    for i in range(levels):
        names += ["downs.%d.0" % i, "downs.%d.1" % i]
    names += ["mid_block1", "mid_block2"]
    for i in range(levels):
        names += ["ups.%d.0" % i, "ups.%d.1" % i]
    names.append("final_res_block")
    return names


def _pack_trunk(w, sd, kind, groups):
    levels = _levels(sd)
    dim = sd["init_conv.weight"].shape[0]
    mults = [sd["downs.%d.3.weight" % i].shape[0] // dim for i in range(levels)]
    w.add("meta", torch.tensor([kind, dim, groups, levels] + mults, dtype=torch.float32), "f32")
    w.add("init_conv.weight", sd["init_conv.weight"].reshape(dim, -1), "f32")
    w.add("init_conv.bias", sd["init_conv.bias"], "f32")
    mlp_w, mlp_b = [], []
    for name in resblock_order(levels):
        for blk in ("block1", "block2"):
            p = "%s.%s" % (name, blk)
            ws = standardize(sd[p + ".proj.weight"].float())
            w.add(p + ".proj.weight", conv_weight_kmajor(ws), "f16")
            if blk == "block1" and ws.shape[0] == 64 and ws.shape[1] == 128:
                # source-split copies for the two-pass halo convolution of cat(x, skip)
                w.add(p + ".proj.weight.a", conv_weight_kmajor(ws[:, :64].contiguous()), "f16")
                w.add(p + ".proj.weight.b", conv_weight_kmajor(ws[:, 64:].contiguous()), "f16")
            w.add(p + ".proj.bias", sd[p + ".proj.bias"], "f32")
            w.add(p + ".norm.weight", sd[p + ".norm.weight"], "f32")
            w.add(p + ".norm.bias", sd[p + ".norm.bias"], "f32")
        if (name + ".res_conv.weight") in sd:
            w.add(name + ".res_conv.weight",
                  conv_weight_kmajor(sd[name + ".res_conv.weight"].float()), "f16")
            w.add(name + ".res_conv.bias", sd[name + ".res_conv.bias"], "f32")
        if (name + ".mlp.1.weight") in sd:
            mlp_w.append(sd[name + ".mlp.1.weight"].float())
            mlp_b.append(sd[name + ".mlp.1.bias"].float())
    if mlp_w:
        w.add("mlp_all.weight", torch.cat(mlp_w, 0), "f32")
        w.add("mlp_all.bias", torch.cat(mlp_b, 0), "f32")
    attn = ["downs.%d.2" % i for i in range(levels)] + ["ups.%d.2" % i for i in range(levels)]
    for p in attn:
        c = sd[p + ".fn.norm.g"].numel()
        w.add(p + ".fn.norm.g", sd[p + ".fn.norm.g"].reshape(c), "f32")
        w.add(p + ".fn.fn.to_qkv.weight", conv_weight_kmajor(sd[p + ".fn.fn.to_qkv.weight"].float()),
              "f16")
        # fp32: folded with the per-image context into W_eff on the device
        w.add(p + ".fn.fn.to_out.0.weight", sd[p + ".fn.fn.to_out.0.weight"].reshape(c, -1), "f32")
        w.add(p + ".fn.fn.to_out.0.bias", sd[p + ".fn.fn.to_out.0.bias"], "f32")
        w.add(p + ".fn.fn.to_out.1.g", sd[p + ".fn.fn.to_out.1.g"].reshape(c), "f32")
    p = "mid_attn"
    c = sd[p + ".fn.norm.g"].numel()
    w.add(p + ".fn.norm.g", sd[p + ".fn.norm.g"].reshape(c), "f32")
    w.add(p + ".fn.fn.to_qkv.weight", conv_weight_kmajor(sd[p + ".fn.fn.to_qkv.weight"].float()), "f16")
    w.add(p + ".fn.fn.to_out.weight", conv_weight_kmajor(sd[p + ".fn.fn.to_out.weight"].float()), "f16")
    w.add(p + ".fn.fn.to_out.bias", sd[p + ".fn.fn.to_out.bias"], "f32")
    for i in range(levels):
        p = "downs.%d.3" % i
        w.add(p + ".weight", conv_weight_kmajor(sd[p + ".weight"].float()), "f16")
        w.add(p + ".bias", sd[p + ".bias"], "f32")
        p = "ups.%d.3" % i
        if (p + ".1.weight") in sd:
            w.add(p + ".1.weight", upsample_fold_weight(sd[p + ".1.weight"].float()), "f16")
            if tuple(sd[p + ".1.weight"].shape[:2]) == (64, 128):
                w.add(p + ".1.weight.rows3", upsample_rows3_weight(sd[p + ".1.weight"].float()), "f16")
            w.add(p + ".1.bias", sd[p + ".1.bias"], "f32")
        else:
            w.add(p + ".weight", conv_weight_kmajor(sd[p + ".weight"].float()), "f16")
            w.add(p + ".bias", sd[p + ".bias"], "f32")


def pack_unet(sd, groups=8):
    """Unet state-dict (280 entries, SDD:802-918) -> blob bytes."""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    w = BlobWriter()
    _pack_trunk(w, sd, KIND_UNET, groups)
    for k in ("time_mlp.1", "time_mlp.3", "param_mlp.0", "param_mlp.2"):
        w.add(k + ".weight", sd[k + ".weight"], "f32")
        w.add(k + ".bias", sd[k + ".bias"], "f32")
    w.add("final_conv.weight", sd["final_conv.weight"].reshape(-1), "f32")
    w.add("final_conv.bias", sd["final_conv.bias"], "f32")
    return w.tobytes()


def pack_maskunet(sd, groups=8):
    """MaskUnet state-dict (234 entries, DC:807-869) -> blob bytes."""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    w = BlobWriter()
    _pack_trunk(w, sd, KIND_MASKUNET, groups)
    w.add("final_conv.0.weight", sd["final_conv.0.weight"].reshape(-1), "f32")
    w.add("final_conv.0.bias", sd["final_conv.0.bias"], "f32")
    return w.tobytes()
