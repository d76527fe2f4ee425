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
    vox = c[c.index("constexpr unsigned long long kVoxEmpty"):c.index("}  // namespace prg")]
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "emu.cpp")
        open(src, "w").write('#include "cuda_shim.h"\n' + occ + driver("DRIVER") + vox + driver("VOX_DRIVER") + MAIN)
        exe = os.path.join(d, "emu")
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                               "-ffp-contract=off", "-I", os.path.join(ROOT, "tests", "emu"), "-o", exe, src])
        return subprocess.call([exe])


if __name__ == "__main__":
    sys.exit(main())
