// sample_host.cpp -- test harness: the record arithmetic of the front end (llcomp_b200/csrc/sample.cuh) compiled for the
// HOST and checked exhaustively against the plain packing of common.cuh (tests/test_sample_host.py).  Not part of the
// product.
#include <algorithm>
#include <cstdint>
using std::max;
using std::min;
#include "../../llcomp_b200/csrc/common.cuh"
#include "../../llcomp_b200/csrc/sample.cuh"

// record_of(hash, diff) must be pack_symbol(|hash|, hash < 0 ? -diff : diff) (llcomp.hpp:431-436) for every context hash
// and every residual a pair of plane values can give.  Returns the number of mismatches.
extern "C" long long sample_check_record_of() {
    long long bad = 0;
    for (int hash = -7925; hash <= 7925; ++hash)
        for (int diff = -1023; diff <= 1023; ++diff) {
            const uint32_t want = llc::pack_symbol(hash < 0 ? -hash : hash, hash < 0 ? -diff : diff);
            bad += llc::record_of(hash, diff) != want;
        }
    return bad;
}

// code_sample (five look-ups) for given neighbours; used to compare with the oracle's front end on the host.
extern "C" uint32_t sample_code(int cur, int l, int L, int tl, int t, int tr, int T) {
    static const llc::QuantBytes q = llc::make_quant_bytes();
    return llc::code_sample(cur, l, L, tl, t, tr, 3025 * q.q5[llc::kQB + T - t], q.q11 + llc::kQB, q.q5 + llc::kQB);
}
