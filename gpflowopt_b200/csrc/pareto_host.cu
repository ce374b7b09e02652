// Host-side native code of the path's setup: the cell decomposition of the non-dominated region that feeds the HVPoI
// kernel (gpflowopt/pareto.py:173-236, Pareto.divide_conquer_nd; pure-Python stack loop in the reference: ~5 s for a
// 200-point 3-objective front, i.e. 20x the time the GPU needs to score 1e6 candidates against the result).
// Same algorithm, same visiting order (LIFO stack, split of the widest edge, upper half pushed last), same acceptance /
// rejection tests with stability = 1e-6, so the int32 tables are identical to the reference's row for row.
#include <cmath>
#include <cstring>
#include <vector>
#include "gpfo_internal.cuh"

extern "C" int gpfo_pareto_cells(const double* front, const int32_t* order, int P, int R, double threshold, int32_t* lb,
                                 int32_t* ub, int64_t capacity, int64_t* count) {
    if (!front || !order || !count || P < 1 || R < 2 || R > 16 || capacity < 0) return GPFO_ERR_BAD_SHAPE;
    const double stability = 1e-6;
    // extended front: row 0 = min - 1 (ideal), rows 1..P = front, row P+1 = max + 1 (anti-ideal)
    std::vector<double> ext((size_t)(P + 2) * R);
    std::vector<int32_t> rows((size_t)(P + 2) * R);          // rows[k][j]: row of ext holding the k-th smallest value of column j
    double total = 1.0;
    for (int j = 0; j < R; ++j) {
        double lo = front[j], hi = front[j];
        for (int p = 1; p < P; ++p) { lo = std::fmin(lo, front[(size_t)p * R + j]); hi = std::fmax(hi, front[(size_t)p * R + j]); }
        ext[j] = lo - 1.0;
        ext[(size_t)(P + 1) * R + j] = hi + 1.0;
        rows[j] = 0;
        rows[(size_t)(P + 1) * R + j] = P + 1;
        for (int p = 0; p < P; ++p) {
            ext[(size_t)(p + 1) * R + j] = front[(size_t)p * R + j];
            rows[(size_t)(p + 1) * R + j] = order[(size_t)p * R + j] + 1;
        }
    }
    for (int j = 0; j < R; ++j) total *= ext[(size_t)(P + 1) * R + j] - ext[j];

    // "the test point augments or dominates the front": smaller than every front member in at least one objective
    auto outside = [&](const double* corner) {
        for (int p = 0; p < P; ++p) {
            bool any = false;
            for (int j = 0; j < R && !any; ++j) any = corner[j] < front[(size_t)p * R + j];
            if (!any) return false;
        }
        return true;
    };

    struct Cell { int32_t a[16], b[16]; };
    std::vector<Cell> stack;
    Cell first;
    for (int j = 0; j < R; ++j) { first.a[j] = 0; first.b[j] = P + 1; }
    stack.push_back(first);
    std::vector<int32_t> out_lb, out_ub;
    double lbv[16], ubv[16], corner[16];
    int32_t lbr[16], ubr[16];
    while (!stack.empty()) {
        const Cell cell = stack.back();
        stack.pop_back();
        for (int j = 0; j < R; ++j) {
            lbr[j] = rows[(size_t)cell.a[j] * R + j];
            ubr[j] = rows[(size_t)cell.b[j] * R + j];
            lbv[j] = ext[(size_t)lbr[j] * R + j];
            ubv[j] = ext[(size_t)ubr[j] * R + j];
        }
        for (int j = 0; j < R; ++j) corner[j] = ubv[j] - stability;
        if (outside(corner)) {                                   // valid integration cell: keep
            out_lb.insert(out_lb.end(), lbr, lbr + R);
            out_ub.insert(out_ub.end(), ubr, ubr + R);
            continue;
        }
        for (int j = 0; j < R; ++j) corner[j] = lbv[j] + stability;
        if (!outside(corner)) continue;                          // lower corner dominated: discard
        int32_t width = 0, k = 0;
        bool any_wide = false;
        for (int j = 0; j < R; ++j) {
            const int32_t d = cell.b[j] - cell.a[j];
            if (d > 1) any_wide = true;
            if (d > width) { width = d; k = j; }                 // np.argmax: first maximum
        }
        double vol = 1.0;
        for (int j = 0; j < R; ++j) vol *= ubv[j] - lbv[j];
        if (any_wide && (vol / total) > threshold) {
            const int32_t half = (int32_t)std::nearbyint(width / 2.0);   // np.round: half to even
            Cell lower = cell, upper = cell;
            lower.b[k] -= half;
            upper.a[k] += width - half;
            stack.push_back(lower);
            stack.push_back(upper);
        }
    }
    const int64_t C = (int64_t)(out_lb.size() / R);
    *count = C;
    if (C > capacity) return GPFO_ERR_BAD_SHAPE;                 // caller retries with a larger buffer
    if (C > 0 && (!lb || !ub)) return GPFO_ERR_BAD_SHAPE;
    if (C > 0) {
        std::memcpy(lb, out_lb.data(), out_lb.size() * sizeof(int32_t));
        std::memcpy(ub, out_ub.data(), out_ub.size() * sizeof(int32_t));
    }
    return GPFO_OK;
}

// Non-dominated sort of the reference (gpflowopt/pareto.py:78-90): dominance[i] = number of rows k with
// all_j Y[k][j] <= Y[i][j] and any_j Y[k][j] < Y[i][j]; the front is {i : dominance[i] == 0}.  The reference builds an
// n x n x R boolean temporary; the NumPy restatement loops over rows (70 ms at n = 1024, 3.6 s at n = 8192 -- more than the
// device spends scoring 1e6 candidates against the result).  O(n^2 R) here too, but branch-light C++: 1 ms / 70 ms.
// Comparisons with NaN are false on both sides, as in NumPy.
extern "C" int gpfo_non_dominated_sort(const double* Y, int64_t n, int R, int64_t* dominance) {
    if (!Y || !dominance || n < 0 || R < 1 || R > 64) return GPFO_ERR_BAD_SHAPE;
    // column-major copy: the inner loop over i then runs over contiguous doubles and vectorises (no short-circuit branches)
    std::vector<double> col((size_t)n * R);
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < R; ++j) col[(size_t)j * n + i] = Y[(size_t)i * R + j];
    std::vector<unsigned char> no_worse((size_t)n), better((size_t)n);
    for (int64_t i = 0; i < n; ++i) dominance[i] = 0;
    for (int64_t k = 0; k < n; ++k) {
        for (int64_t i = 0; i < n; ++i) { no_worse[i] = 1; better[i] = 0; }
        for (int j = 0; j < R; ++j) {
            const double ykj = Y[(size_t)k * R + j];
            const double* __restrict__ cj = col.data() + (size_t)j * n;
            unsigned char* __restrict__ nw = no_worse.data();
            unsigned char* __restrict__ bt = better.data();
            for (int64_t i = 0; i < n; ++i) {
                nw[i] &= (unsigned char)(ykj <= cj[i]);
                bt[i] |= (unsigned char)(ykj < cj[i]);
            }
        }
        for (int64_t i = 0; i < n; ++i) dominance[i] += no_worse[i] & better[i];
    }
    return GPFO_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Candidate rows of RandomDesign (gpflowopt/design.py:55-70, 83-94) for the multi-GPU Monte-Carlo route, where rank 0 alone
// draws the reference's single np.random stream: n x D doubles of np.random.rand -- MT19937, 53-bit doubles
// (a >> 5, b >> 6 of two consecutive 32-bit outputs: (a * 2^26 + b) / 2^53) -- continued from the caller's RandomState
// (key[624], pos as np.random.get_state() returns them; both are updated so that np.random.set_state resumes the stream),
// with the design's diagonal map x * scale[d] + shift[d] applied in the same pass.  Bit-identical to
// np.random.rand(n, D) * scale + shift; about 4 x faster than NumPy's generator plus two more passes over the array, which
// at 1e7 x 16 is what bounds the wall time of BASELINE configs[2] on 8 GPUs.
// ------------------------------------------------------------------------------------------------------------------------
static inline void mt19937_refill(uint32_t* mt) {
    const int N = 624, Mm = 397;
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    int kk = 0;
    for (; kk < N - Mm; ++kk) {
        const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + Mm] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    for (; kk < N - 1; ++kk) {
        const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (Mm - N)] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    const uint32_t y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
    mt[N - 1] = mt[Mm - 1] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
}
static inline uint32_t mt19937_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
extern "C" int gpfo_mt19937_design_rows(uint32_t* key, int32_t* pos, int64_t n, int D, const double* scale, const double* shift,
                                        double* out) {
    if (!key || !pos || !out || n < 0 || D < 1 || *pos < 0 || *pos > 624) return GPFO_ERR_BAD_SHAPE;
    int p = *pos;
    const int64_t total = n * D;
    int d = 0;
    for (int64_t e = 0; e < total; ++e) {
        if (p >= 624) { mt19937_refill(key); p = 0; }
        const uint32_t a = mt19937_temper(key[p++]) >> 5;
        if (p >= 624) { mt19937_refill(key); p = 0; }
        const uint32_t b = mt19937_temper(key[p++]) >> 6;
        double v = (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);   // division by 2^53 == multiplication by 2^-53, exactly
        if (scale) v = v * scale[d];                 // separate roundings, as NumPy's multiply then add (this translation unit is
        if (shift) v = v + shift[d];                 // host code compiled for baseline x86-64: no FMA contraction)
        out[e] = v;
        if (++d == D) d = 0;
    }
    *pos = p;
    return GPFO_OK;
}
