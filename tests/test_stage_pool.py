"""CPU: the fork-join copy pool behind the staged (pageable-memory) host path of pp_dubins_eval
(rs-pathplanning_b200/csrc/pp_stage.hpp) -- plain C++, so it is compiled with g++ and driven here: every byte of
every piece lands, shares never overlap or run past a piece, odd sizes / empty pieces / one thread all work."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pp_stage.hpp"
int main() {
    const size_t sizes[] = {0, 1, 63, 64, 65, 1000, 4096, 65537, (size_t)3 << 20, ((size_t)5 << 20) + 13};
    for (int threads : {1, 2, 3, 4, 7, 16}) {
        pp_stage_pool pool(threads);
        if (pool.threads() != threads) return 2;
        for (int rep = 0; rep < 20; ++rep) {  // many runs on one pool: the generation hand-shake must not lose a job
            std::vector<std::vector<unsigned char>> src, dst;
            std::vector<pp_copy_piece> pieces;
            unsigned seed = 12345u + rep;
            for (size_t sz : sizes) {
                src.emplace_back(sz);
                dst.emplace_back(sz + 128, 0xEE);  // 64 guard bytes on either side
                for (size_t i = 0; i < sz; ++i) src.back()[i] = (unsigned char)((seed = seed * 1664525u + 1013904223u) >> 24);
            }
            for (size_t k = 0; k < src.size(); ++k) pieces.push_back({dst[k].data() + 64, src[k].data(), src[k].size()});
            pool.run(pieces.data(), pieces.size());
            for (size_t k = 0; k < src.size(); ++k) {
                for (size_t i = 0; i < 64; ++i)
                    if (dst[k][i] != 0xEE || dst[k][64 + src[k].size() + i] != 0xEE) return 3;  // wrote outside the piece
                if (memcmp(dst[k].data() + 64, src[k].data(), src[k].size()) != 0) return 4;    // bytes missing / wrong
            }
        }
        pool.run(nullptr, 0);
    }
    puts("stage pool ok");
    return 0;
}
"""


def test_stage_pool_copies_every_byte(tmp_path):
    src = tmp_path / "stage_harness.cpp"
    src.write_text(HARNESS)
    exe = str(tmp_path / "stage_harness")
    inc = os.path.join(ROOT, "rs-pathplanning_b200", "csrc")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread", "-I", inc, str(src), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "stage pool ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
