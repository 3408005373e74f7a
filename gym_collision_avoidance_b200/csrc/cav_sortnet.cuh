// cav_sortnet.cuh -- compare-exchange schedules for sorting a small compile-time number of register-resident keys.
// Minimal-size networks for n <= 9 and n = 11 (each checked exhaustively with the 0-1 principle by
// tests/test_sortnet_cpu.py, which parses this file); any other n takes odd-even transposition (n rounds).
// `cex(i, j)` must leave the smaller element at i (i < j); the indices are literals, so arrays stay in registers.
#pragma once
#include "cav_hd.cuh"

namespace cav {

template <typename F>
CAV_HDI void sortnet_2(F &&cex) {  // 1 compare-exchanges
    cex(0, 1);
}
template <typename F>
CAV_HDI void sortnet_3(F &&cex) {  // 3 compare-exchanges
    cex(0, 2); cex(0, 1); cex(1, 2);
}
template <typename F>
CAV_HDI void sortnet_4(F &&cex) {  // 5 compare-exchanges
    cex(0, 2); cex(1, 3); cex(0, 1); cex(2, 3); cex(1, 2);
}
template <typename F>
CAV_HDI void sortnet_5(F &&cex) {  // 9 compare-exchanges
    cex(0, 3); cex(1, 4); cex(0, 2); cex(1, 3); cex(0, 1); cex(2, 4); cex(1, 2); cex(3, 4); cex(2, 3);
}
template <typename F>
CAV_HDI void sortnet_6(F &&cex) {  // 12 compare-exchanges
    cex(0, 5); cex(1, 3); cex(2, 4); cex(1, 2); cex(3, 4); cex(0, 3); cex(2, 5); cex(0, 1); cex(2, 3); cex(4, 5);
    cex(1, 2); cex(3, 4);
}
template <typename F>
CAV_HDI void sortnet_7(F &&cex) {  // 16 compare-exchanges
    cex(0, 6); cex(2, 3); cex(4, 5); cex(0, 2); cex(1, 4); cex(3, 6); cex(0, 1); cex(2, 5); cex(3, 4); cex(1, 2);
    cex(4, 6); cex(2, 3); cex(4, 5); cex(1, 2); cex(3, 4); cex(5, 6);
}
template <typename F>
CAV_HDI void sortnet_8(F &&cex) {  // 19 compare-exchanges
    cex(0, 2); cex(1, 3); cex(4, 6); cex(5, 7); cex(0, 4); cex(1, 5); cex(2, 6); cex(3, 7); cex(0, 1); cex(2, 3);
    cex(4, 5); cex(6, 7); cex(2, 4); cex(3, 5); cex(1, 4); cex(3, 6); cex(1, 2); cex(3, 4); cex(5, 6);
}
template <typename F>
CAV_HDI void sortnet_9(F &&cex) {  // 25 compare-exchanges
    cex(0, 3); cex(1, 7); cex(2, 5); cex(4, 8); cex(0, 7); cex(2, 4); cex(3, 8); cex(5, 6); cex(0, 2); cex(1, 3);
    cex(4, 5); cex(7, 8); cex(1, 4); cex(3, 6); cex(5, 7); cex(0, 1); cex(2, 4); cex(3, 5); cex(6, 8); cex(2, 3);
    cex(4, 5); cex(6, 7); cex(1, 2); cex(3, 4); cex(5, 6);
}
template <typename F>
CAV_HDI void sortnet_11(F &&cex) {  // 35 compare-exchanges
    cex(0, 9); cex(1, 6); cex(2, 4); cex(3, 7); cex(5, 8); cex(0, 1); cex(3, 5); cex(4, 10); cex(6, 9); cex(7, 8);
    cex(1, 3); cex(2, 5); cex(4, 7); cex(8, 10); cex(0, 4); cex(1, 2); cex(3, 7); cex(5, 9); cex(6, 8); cex(0, 1);
    cex(2, 6); cex(4, 5); cex(7, 8); cex(9, 10); cex(2, 4); cex(3, 6); cex(5, 7); cex(8, 9); cex(1, 2); cex(3, 4);
    cex(5, 6); cex(7, 8); cex(2, 3); cex(4, 5); cex(6, 7);
}

template <int N, typename F>
CAV_HDI void sort_network(F &&cex) {
    if constexpr (N < 2) {
    }
    else if constexpr (N == 2) sortnet_2(cex);
    else if constexpr (N == 3) sortnet_3(cex);
    else if constexpr (N == 4) sortnet_4(cex);
    else if constexpr (N == 5) sortnet_5(cex);
    else if constexpr (N == 6) sortnet_6(cex);
    else if constexpr (N == 7) sortnet_7(cex);
    else if constexpr (N == 8) sortnet_8(cex);
    else if constexpr (N == 9) sortnet_9(cex);
    else if constexpr (N == 11) sortnet_11(cex);
    else {
#pragma unroll
        for (int r = 0; r < N; ++r) {
#pragma unroll
            for (int i = (r & 1); i + 1 < N; i += 2) cex(i, i + 1);
        }
    }
}

}  // namespace cav
