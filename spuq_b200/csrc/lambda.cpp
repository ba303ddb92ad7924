// lambda.cpp -- K5: integer set-up of the active multi-index set (host, bit-exact).
//
// Replaces, for the whole set at once, what the reference does index by index in Python:
//   * ordering   Multiindex.cmp_by_order, math_utils/multiindex.py:42-53  (graded, then entries
//                left to right, lower entry first) as used by sorted() in
//                MultiVector.active_indices(), application/egsz/multi_vector.py:97-98
//   * lookups    `mu.inc(m) in Lambda`, `mu.dec(m) in Lambda`,
//                application/egsz/multi_operator2.py:72-79, which scan a Python list with
//                sha1-based equality (O(L) each); here one open-addressing hash table over the
//                rows, O(1) per lookup.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "common.h"

namespace {

inline uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
}

uint64_t hash_row(const int32_t* r, int32_t M) {
    // trailing zeros do not change the identity of a multi-index (multiindex.py:26-32), and all
    // rows have the same width M here, so hashing the full padded row is consistent.
    uint64_t h = 0x1234567u;
    for (int32_t j = 0; j < M; ++j) h = mix(h, (uint64_t)(uint32_t)r[j] * 0x100000001b3ULL + j);
    return h;
}

struct RowTable {
    const int32_t* rows;
    int32_t L, M;
    std::vector<int32_t> slot;
    uint64_t mask;

    RowTable(const int32_t* r, int32_t L_, int32_t M_) : rows(r), L(L_), M(M_) {
        uint64_t cap = 16;
        while (cap < (uint64_t)L * 2 + 2) cap <<= 1;
        mask = cap - 1;
        slot.assign(cap, -1);
    }
    // returns existing position or -1 after inserting k
    int32_t insert(int32_t k) {
        const int32_t* r = rows + (int64_t)k * M;
        uint64_t h = hash_row(r, M) & mask;
        while (slot[h] >= 0) {
            if (std::memcmp(rows + (int64_t)slot[h] * M, r, sizeof(int32_t) * M) == 0) return slot[h];
            h = (h + 1) & mask;
        }
        slot[h] = k;
        return -1;
    }
    int32_t find(const int32_t* r) const {
        uint64_t h = hash_row(r, M) & mask;
        while (slot[h] >= 0) {
            if (std::memcmp(rows + (int64_t)slot[h] * M, r, sizeof(int32_t) * M) == 0) return slot[h];
            h = (h + 1) & mask;
        }
        return -1;
    }
};

}  // namespace

extern "C" int spuq_sg_lambda_sort(const int32_t* mi, int32_t L, int32_t M, int32_t* perm) {
    SPUQ_REQUIRE(L >= 0 && M >= 0, "lambda_sort: negative size");
    SPUQ_REQUIRE(L == 0 || (mi && perm), "lambda_sort: null pointer");
    std::vector<int64_t> deg(L, 0);
    for (int32_t k = 0; k < L; ++k) {
        for (int32_t j = 0; j < M; ++j) {
            int32_t v = mi[(int64_t)k * M + j];
            SPUQ_REQUIRE(v >= 0, "lambda_sort: negative entry in multi-index %d", k);
            deg[k] += v;
        }
    }
    std::iota(perm, perm + L, 0);
    // graded, then lexicographic on the (zero-padded) entries; padding with zeros is equivalent to
    // the reference's comparison over the common prefix because two indices of equal degree that
    // agree on the shorter one's prefix are identical.
    std::stable_sort(perm, perm + L, [&](int32_t a, int32_t b) {
        if (deg[a] != deg[b]) return deg[a] < deg[b];
        const int32_t* ra = mi + (int64_t)a * M;
        const int32_t* rb = mi + (int64_t)b * M;
        for (int32_t j = 0; j < M; ++j)
            if (ra[j] != rb[j]) return ra[j] < rb[j];
        return false;
    });
    for (int32_t k = 1; k < L; ++k) {
        const int32_t* ra = mi + (int64_t)perm[k - 1] * M;
        const int32_t* rb = mi + (int64_t)perm[k] * M;
        SPUQ_REQUIRE(M == 0 ? false : std::memcmp(ra, rb, sizeof(int32_t) * M) != 0,
                     "lambda_sort: duplicate multi-index (rows %d and %d)", perm[k - 1], perm[k]);
    }
    return SPUQ_OK;
}

extern "C" int spuq_sg_lambda_neighbours(const int32_t* mi, int32_t L, int32_t M,
                                         int32_t* nbr_up, int32_t* nbr_dn) {
    SPUQ_REQUIRE(L >= 0 && M >= 0, "lambda_neighbours: negative size");
    if (L == 0 || M == 0) return SPUQ_OK;
    SPUQ_REQUIRE(mi && nbr_up && nbr_dn, "lambda_neighbours: null pointer");
    RowTable table(mi, L, M);
    for (int32_t k = 0; k < L; ++k) {
        int32_t prev = table.insert(k);
        SPUQ_REQUIRE(prev < 0, "lambda_neighbours: duplicate multi-index (rows %d and %d)", prev, k);
    }
    std::vector<int32_t> probe(M);
    for (int32_t k = 0; k < L; ++k) {
        const int32_t* r = mi + (int64_t)k * M;
        std::memcpy(probe.data(), r, sizeof(int32_t) * M);
        for (int32_t m = 0; m < M; ++m) {
            probe[m] = r[m] + 1;
            nbr_up[(int64_t)m * L + k] = table.find(probe.data());
            if (r[m] > 0) {
                probe[m] = r[m] - 1;
                nbr_dn[(int64_t)m * L + k] = table.find(probe.data());
            } else {
                nbr_dn[(int64_t)m * L + k] = -1;   // Multiindex.dec returns None, multiindex.py:107-108
            }
            probe[m] = r[m];
        }
    }
    return SPUQ_OK;
}
