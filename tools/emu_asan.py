"""Runs the host emulation of the simple kernels (tests/test_kernel_emulation.py) as a stand-alone
program under AddressSanitizer / UBSan with exact-size heap buffers, so that an out-of-bounds load
or store in a kernel's index arithmetic shows up without a GPU.  CPU only.

    python tools/emu_asan.py
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MAIN = r'''
#include <vector>
#include <random>
#include <cstdio>
int main() {
  std::mt19937 rng(1);
  int shapes[][3] = {{2,256,256},{1,480,640},{2,33,47},{1,5,7},{1,40,260},{1,1,1},{1,70,4},{3,17,1},{1,2,1023}};
  for (auto& s : shapes) {
    int B = s[0], H = s[1], W = s[2];
    size_t n = (size_t)B * H * W;
    std::vector<float> d(n), o(n);
    std::vector<uint8_t> m(n);
    for (size_t i = 0; i < n; ++i) { d[i] = (rng() % 1000) * 0.01f; m[i] = rng() % 3 != 0; if (!m[i]) d[i] = 0; }
    emu_occlusion(d.data(), m.data(), o.data(), B, H, W);
  }
  long long ns[] = {1, 3, 100, 4097, 20000};
  for (long long n : ns) {
    std::vector<double> p(n * 3), q(n * 3), cent(n * 3);
    std::vector<long long> keys(n);
    for (auto& v : p) v = (rng() % 100000) * 2e-5 - 1.0;
    for (auto& v : q) v = (rng() % 100000) * 2e-5 - 0.7;
    unsigned long long cap = 1024;
    while (cap < 2ull * n) cap <<= 1;
    std::vector<unsigned char> ws(32 + cap * 36), ws2(cap * 12 + n * 4);
    int ce[2];
    emu_voxel(p.data(), n, 0.025, cent.data(), keys.data(), ce, ws.data(), 7);
    int voxels = ce[0];
    emu_overlap(q.data(), n, p.data(), n, 0.0375, ce, ws2.data(), 5);
    printf("n=%lld: %d voxels, %d overlapping points\n", n, voxels, ce[0]);
  }
  // fused reprojection: exact-size heap buffers for depth, outputs and the ring; odd sizes exercise the
  // partial chunk / partial sub-block paths, item sizes of one to three chunks
  int rshapes[][3] = {{3,256,256},{2,480,640},{2,33,47},{1,5,7},{2,64,96},{1,1,1},{1,90,171},{2,121,127},{1,130,259}};
  int trial = 0;
  for (auto& s : rshapes) {
    int B = s[0], H = s[1], W = s[2];
    size_t n = (size_t)B * H * W;
    for (int ring = 2; ring <= 4; ring += 2, ++trial) {
      std::vector<float> d(n), o(n), K(B * 9, 0.f), P(B * 16, 0.f);
      std::vector<uint8_t> m(n);
      std::vector<unsigned> scratch((size_t)ring * H * W);
      for (size_t i = 0; i < n; ++i) d[i] = (rng() % 7 == 0) ? 0.f : 0.5f + (rng() % 9000) * 0.001f;
      for (int b = 0; b < B; ++b) {
        K[b * 9 + 0] = 1.2f * W; K[b * 9 + 4] = 1.2f * W; K[b * 9 + 2] = 0.5f * W; K[b * 9 + 5] = 0.5f * H; K[b * 9 + 8] = 1.f;
        for (int i = 0; i < 4; ++i) P[b * 16 + i * 5] = 1.f;
        P[b * 16 + 3] = 0.05f * (b + 1); P[b * 16 + 7] = -0.03f; P[b * 16 + 11] = 0.1f;
      }
      emu_reproject(d.data(), K.data(), P.data(), 0.f, 10.f, o.data(), m.data(), scratch.data(), B, H, W, ring, 1 + trial % 3);
      size_t hit = 0;
      for (size_t i = 0; i < n; ++i) hit += m[i];
      if (o[0] == -12345.f) { printf("reproject %dx%dx%d ring %d: ring not handed back empty\n", B, H, W, ring); return 1; }
      printf("reproject %dx%dx%d ring %d item chunks %d: %zu pixels hit\n", B, H, W, ring, 1 + trial % 3, hit);
    }
  }
  puts("asan run complete");
  return 0;
}
'''


def main():
    g = open(os.path.join(ROOT, "pointreggpt_b200", "csrc", "geometry.cu")).read()
    c = open(os.path.join(ROOT, "pointreggpt_b200", "csrc", "cloud.cu")).read()
    t = open(os.path.join(ROOT, "tests", "test_kernel_emulation.py")).read()

    def driver(name):
        return re.search(r"^" + name + r" = '''(.*?)'''", t, re.S | re.M).group(1)

    occ = g[g.index("constexpr int kOccRows"):g.index("// ------------------------------------------------------------------ point_cloud")]
    geo = g[g.index("constexpr unsigned kEmpty"):g.index("// ------------------------------------------------------------------ occlusion_filter")]
    geo = geo.replace("extern __shared__ int64_t s_off[];", "static int64_t s_off[4097];")
    vox = c[c.index("constexpr unsigned long long kVoxEmpty"):c.index("}  // namespace prg")]
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "emu.cpp")
        open(src, "w").write('#include "cuda_shim.h"\n' + geo + driver("GEOM_DRIVER") + occ + driver("DRIVER") + vox +
                             driver("VOX_DRIVER") + MAIN)
        exe = os.path.join(d, "emu")
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                               "-ffp-contract=off", "-I", os.path.join(ROOT, "tests", "emu"), "-o", exe, src])
        return subprocess.call([exe])


if __name__ == "__main__":
    sys.exit(main())
