// chain_host.cu -- test harness: the decoder's slice chain (llcomp_b200/csrc/decoder_chain.cuh) compiled for the HOST,
// so that its logic can be compared with the oracle without a GPU (tests/test_decoder_chain_host.py).  Not part of the
// product: the library only ever runs this code on the device.
#include <cstdint>
#include <cstring>
#include <vector>

#include <algorithm>
using std::min;
using std::max;
#include "../../llcomp_b200/csrc/common.cuh"
#include "../../llcomp_b200/csrc/decoder_chain.cuh"

using namespace llc;
using namespace llc::dchain;

template <int CT>
static int decode(const uint8_t* payload, uint32_t len, int w, int h, uint8_t* out, size_t pitch) {
    static const ModelTables tables = make_tables();
    std::vector<uint8_t> arena(layout_bytes(w * CT), 0xA5);                     // garbage, as shared memory is
    std::vector<Row> rows(kContexts, Row{0, 0});
    Smem m{arena.data()};
    StateMem<true> st{rows.data()};
    const bool ok = decode_slice<CT, true>(m, st, tables.entry, payload, len, w, h, out, pitch, 0, 1, [] {});
    return ok ? 0 : 2;
}

extern "C" int chain_decode_tile(const uint8_t* payload, uint32_t len, int w, int h, int c, uint8_t* out, size_t pitch) {
    switch (c) {
        case 1: return decode<1>(payload, len, w, h, out, pitch);
        case 2: return decode<2>(payload, len, w, h, out, pitch);
        case 3: return decode<3>(payload, len, w, h, out, pitch);
        case 4: return decode<4>(payload, len, w, h, out, pitch);
    }
    return 1;
}
