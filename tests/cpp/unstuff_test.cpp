// unstuff_test.cpp -- the three implementations of ent_unstuff (csrc/entropy_host.h: scalar memchr + memcpy, AVX2,
// AVX-512 VBMI2) against a byte-at-a-time restatement of what they must do: copy bytes, turn FF 00 into FF, stop at
// the first FF followed by anything else and say where it is.  Random buffers with every density of FF / 00 / markers,
// every alignment and length around the vector sizes, destinations that are just large enough.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../jpeg_decoder_b200/csrc/entropy_host.h"

using namespace b200jpg;

// returns the offset of the marker FF (or -1) and the bytes written
static long reference(const std::vector<uint8_t>& in, std::vector<uint8_t>* out) {
    out->clear();
    for (size_t i = 0; i < in.size(); i++) {
        if (in[i] != 0xFF) { out->push_back(in[i]); continue; }
        if (i + 1 >= in.size()) return -1;          // FF at the very end: no pair
        if (in[i + 1] == 0x00) { out->push_back(0xFF); i++; continue; }
        return (long)i;
    }
    return -1;
}

int main() {
    unsigned long long seed = 12345;
    auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(seed >> 33); };
    unsigned checked = 0;
    for (int iter = 0; iter < 20000; iter++) {
        const size_t n = rnd() % 400;
        const unsigned p_ff = rnd() % 5 == 0 ? 2 : (rnd() % 64 + 2), p_marker = rnd() % 3 == 0 ? 1000000 : (rnd() % 40 + 1);
        std::vector<uint8_t> in(n);
        for (size_t i = 0; i < n; i++) {
            in[i] = (uint8_t)rnd();
            if (rnd() % p_ff == 0) in[i] = 0xFF;
            else if (i > 0 && in[i - 1] == 0xFF) in[i] = (rnd() % p_marker == 0) ? (uint8_t)(0xD0 + rnd() % 16) : 0x00;
        }
        const size_t lead = rnd() % 67;  // misalign the source
        std::vector<uint8_t> buf(lead + n + 1, 0xAB);
        memcpy(buf.data() + lead, in.data(), n);
        std::vector<uint8_t> want;
        const long where = reference(in, &want);
        for (int level = 0; level <= 2; level++) {
            ent_unstuff_level() = level;
            std::vector<uint8_t> dst(n + 64 + 256 + rnd() % 64, 0xCD);
            uint8_t* o = dst.data() + rnd() % 5;
            uint8_t* const o0 = o;
            const uint8_t* f = ent_unstuff(buf.data() + lead, buf.data() + lead + n, &o, dst.data() + dst.size());
            const long got_where = f ? (long)(f - (buf.data() + lead)) : -1;
            // the contract of the caller: a marker at `where`, or nullptr when the input ends without one
            if (where >= 0 && (size_t)where + 1 < n) {
                if (got_where != where || (size_t)(o - o0) != want.size() || memcmp(o0, want.data(), want.size()) != 0) {
                    printf("MISMATCH level %d iter %d: where %ld vs %ld, len %zu vs %zu\n", level, iter, got_where, where, (size_t)(o - o0), want.size());
                    return 1;
                }
            } else if (f != nullptr && !(where >= 0 && got_where == where)) {
                printf("MISMATCH level %d iter %d: expected no marker, got %ld\n", level, iter, got_where);
                return 1;
            }
            checked++;
        }
    }
    ent_unstuff_level() = -1;
    printf("ok %u cases\n", checked);
    return 0;
}
