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

template <int CT, int kV>
static int decode(const uint8_t* payload, uint32_t len, int w, int h, uint8_t* out, size_t pitch) {
    static const ModelTables tables = make_tables();
    std::vector<uint8_t> arena(layout_bytes(w * CT), 0xA5);                     // garbage, as shared memory is
    std::vector<Row> rows(kContexts, Row{0, 0});
    Smem m{arena.data()};
    StateMem<true> st{rows.data()};
    const bool ok = decode_slice_rows<CT, true, kV>(m, make_layout(w * CT, 0), st, tables.entry, payload, len, w, h, out, pitch, 0, 1, [] {});
    return ok ? 0 : 2;
}

// variant: Chain's kV (0 = default decisions, 4 = products taken ahead of the renormalisation test)
extern "C" int chain_decode_tile_v(const uint8_t* payload, uint32_t len, int w, int h, int c, uint8_t* out, size_t pitch,
                                   int variant) {
#define CASE(C) case C: return variant == 4 ? decode<C, 4>(payload, len, w, h, out, pitch) : decode<C, 0>(payload, len, w, h, out, pitch);
    switch (c) { CASE(1) CASE(2) CASE(3) CASE(4) }
#undef CASE
    return 1;
}
extern "C" int chain_decode_tile(const uint8_t* payload, uint32_t len, int w, int h, int c, uint8_t* out, size_t pitch) {
    return chain_decode_tile_v(payload, len, w, h, c, out, pitch, 0);
}
