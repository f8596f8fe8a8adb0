// Host cost of one raw position (dg_raw_position: stones, hashes, the two ladder planes) on positions from uniformly random
// legal playouts (plies 80-280: weak chains everywhere, ~18 ladder readings per position), with a checksum of the ladder
// planes so that variants of the reader can be compared.
//   g++ -O3 -march=x86-64-v3 -ffp-contract=off -std=c++17 -Idream_go_b200/csrc -Iinclude tools/bench_ladder.cpp -o /tmp/bench_ladder && /tmp/bench_ladder
#include <chrono>
#include <cstdio>
#include <vector>
#include "go_board.h"
#include "dg_engine.h"
using namespace dg;
int main() {
    std::vector<Board> pos; std::vector<int> tm;
    uint64_t rng = 12345;
    auto next = [&] { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); };
    for (int g = 0; g < 40; ++g) {
        Board b; b.init(7.5f); int color = BLACK;
        for (int ply = 0; ply < 280; ++ply) {
            std::vector<int> cand;
            for (int p = 0; p < N_POINTS; ++p) if (b.is_valid(color, p) && !is_simple_eye(b, color, p)) cand.push_back(p);
            if (cand.empty()) break;
            b.place(color, cand[next() % cand.size()]); color = opposite(color);
            if (ply >= 80 && ply % 10 == 0) { pos.push_back(b); tm.push_back(color); }
        }
    }
    printf("%zu positions\n", pos.size());
    dg_raw_position raw;
    uint64_t sum = 0;
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        sum = 0;
        for (int r = 0; r < 20; ++r)
            for (size_t i = 0; i < pos.size(); ++i) {
                raw_position(pos[i], tm[i], 0, &raw);
                for (int k = 0; k < 12; ++k) sum = sum * 1099511628211ull + raw.ladder_capture[k] + 7ull * raw.ladder_escape[k];
            }
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("raw_position %.2f us/position  checksum %llx\n", 1e6 * dt / (20.0 * pos.size()), (unsigned long long)sum);
    }
}
