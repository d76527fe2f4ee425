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
