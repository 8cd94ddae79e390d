// entropy_emul.cpp -- CPU emulation of the device entropy decoder (csrc/entropy_dev.h), pass by pass, with the very
// functions the kernels call (ent_decode_range, EntWriteSink, ent_build_tables, ent_build_payload, ent_fill_image).
// For each JPEG file given:
//   * decodes it with the host decoder (the reference's sequential Huffman loop) -> dense coefficients
//   * runs cold / sync* / prefix / write / dc exactly as the kernels do (CTA-local rounds, predecessor-changed flags)
//   * requires: device result accepted  =>  host decode succeeded and every coefficient is identical
// With --corrupt N SEED the same is repeated for N corrupted copies of each file (random byte edits inside the
// scan): the device path must either flag the image or agree with the host bit for bit.
// Test infrastructure only (tests/test_entropy_emul.py builds it with g++).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "../../jpeg_decoder_b200/csrc/entropy_host.h"

using namespace b200jpg;

static const uint8_t UNZZ[64] = ENT_UNZIGZAG_INIT;

static int g_cta_order = 0;

struct EmulResult {
    bool eligible = false;
    unsigned status = 0;  // anomaly bits
    unsigned passes = 0, nsub = 0, intervals = 0;
    std::vector<int16_t> coefs;  // all components back to back
    std::vector<uint8_t> cs;     // the compact stream of the image (entropy_dev.h)
};

struct EmulResult;
static void emulate_interval(const std::vector<uint8_t>& payload, const EntImage& im, const b200jpg_image_desc& d, int max_passes, EmulResult& r);

static EmulResult emulate(const std::vector<uint8_t>& file, const HostDecoder& probe, const b200jpg_image_desc& d, int max_passes) {
    EmulResult r;
    std::vector<uint8_t> payload(ent_payload_bound(file.size()) + 64);
    const size_t plen = ent_build_payload(probe, file.data(), file.size(), payload.data(), payload.size());
    if (!plen) return r;
    EntHeader h;
    memcpy(&h, payload.data(), sizeof h);
    size_t coef_off[4] = {0, 0, 0, 0}, total = 0;
    for (int c = 0; c < d.ncomp; c++) {
        coef_off[c] = total;
        total += (size_t)d.comps[c].block_w * d.comps[c].block_h * 128;
    }
    std::vector<EntImage> ims(h.nintervals ? h.nintervals : 1);
    unsigned nsub_total = 0;
    if (!ent_fill_images(payload.data(), plen, d, coef_off, 0, 0, 0, ims.data(), ims.size(), &nsub_total)) return r;
    r.eligible = true;
    r.nsub = nsub_total;
    r.intervals = (unsigned)ims.size();
    r.coefs.assign(total / 2, 0x5a5a);  // K0 writes every coefficient, zeros included
    r.cs.assign(ent_cs_bytes(h.total_blocks), 0xee);
    memset(r.cs.data(), 0, ent_cs_header_bytes(ims[0].nb_pad));  // k0_zero_headers
    for (const EntImage& im : ims) emulate_interval(payload, im, d, max_passes, r);
    // K0 (k0_expand.cu) on the compact stream: bitmaps + values -> dense blocks in the slab
    const EntImage& im0 = ims[0];
    const unsigned long long* bm = (const unsigned long long*)r.cs.data();
    const int16_t* dc = (const int16_t*)(r.cs.data() + 8 * (size_t)im0.nb_pad);
    const uint32_t* boff = (const uint32_t*)(r.cs.data() + 10 * (size_t)im0.nb_pad);
    for (unsigned t = 0; t < h.total_blocks; t++) {
        const unsigned m = t / im0.bpm, j = t % im0.bpm, c = im0.mcu_comp[j];
        const unsigned mx = m % im0.mcu_w, my = m / im0.mcu_w;
        int16_t* blk = r.coefs.data() + ((size_t)im0.slab_row[c] + (size_t)(my * im0.v[c] + im0.mcu_vy[j]) * im0.block_w[c] + mx * im0.h[c] + im0.mcu_hx[j]) * 64;
        const int16_t* v = (const int16_t*)(r.cs.data() + boff[t]);
        unsigned rank = 0;
        for (unsigned k = 0; k < 64; k++) {
            if (k == 0) blk[0] = dc[t];
            else if ((bm[t] >> k) & 1) blk[UNZZ[k]] = v[rank++];
            else blk[UNZZ[k]] = 0;
        }
    }
    return r;
}

// ENT_EMUL_ROUNDS=1: per synchronisation launch and CTA-local round, how many CTAs were still iterating and how many
// subsequences they re-decoded (what the latency of ent_sync is made of)
struct RoundStat { unsigned ctas = 0, pending = 0, warps = 0; };
static bool g_round_stats = getenv("ENT_EMUL_ROUNDS") != nullptr;
static std::vector<std::vector<RoundStat>> g_rounds(64);

// one interval (= the whole scan without DRI): what the kernels do for one "image"
static void emulate_interval(const std::vector<uint8_t>& payload, const EntImage& im, const b200jpg_image_desc& d, int max_passes, EmulResult& r) {
    const EntWordsGlobal words{(const uint32_t*)(payload.data() + im.data_off), im.nwords};
    const uint16_t* tabs = (const uint16_t*)(payload.data() + im.tables_off);
    const unsigned n = im.nsub;
    std::vector<uint64_t> state(n);
    std::vector<uint8_t> chA(n, 1), chB(n, 0);
    std::vector<uint32_t> nvals(n, 0);
    unsigned dummy = 0, passes = 0;
    // cold
    for (unsigned i = 0; i < n; i++) {
        EntCountSink sink;
        EntState st{i * ENT_SUB_BITS, 0, 0, 0};
        state[i] = ent_pack(ent_decode_range<false>(words, tabs, im.dcslot, im.acslot, im.dec_bpm, st, ent_sub_end(i, n, im.scan_bits), sink, &dummy));
        nvals[i] = sink.nvals;
    }
    // sync, scheduled like ent_sync (ke_entropy.cu): per launch every CTA of 128 subsequences iterates on its own until
    // it is quiet (flags through "shared memory"), states are read and written in place, only the hand-over between
    // CTAs waits for the next launch.  g_cta_order: 0 = CTAs ascending, 1 = descending (the GPU runs them in any order).
    uint8_t *cin = chA.data(), *cout = chB.data();
    const unsigned T = 256, LOCAL = 32;  // ENT_THREADS, ENT_LOCAL_ITERS of ke_entropy.cu
    for (int pass = 0; pass < max_passes; pass++) {
        unsigned any = 0;
        const unsigned nctas = (n + T - 1) / T;
        for (unsigned cc = 0; cc < nctas; cc++) {
            const unsigned cta = g_cta_order ? nctas - 1 - cc : cc;
            const unsigned i0 = cta * T, cnt = std::min(T, n - i0);
            std::vector<uint8_t> pending(cnt), changed(cnt, 0), ever(cnt, 0);
            bool anyp = false;
            for (unsigned t = 0; t < cnt; t++) {
                pending[t] = (i0 + t) > 0 && cin[i0 + t - 1];
                anyp |= pending[t];
            }
            if (!anyp) {
                for (unsigned t = 0; t < cnt; t++) cout[i0 + t] = 0;
                continue;
            }
            std::vector<uint64_t> mine(state.begin() + i0, state.begin() + i0 + cnt);
            for (unsigned iter = 0; iter < LOCAL; iter++) {
                if (g_round_stats) {
                    unsigned np = 0;
                    for (unsigned t = 0; t < cnt; t++) np += pending[t];
                    auto& st = g_rounds[(size_t)pass];
                    if (st.size() <= iter) st.resize(iter + 1);
                    st[iter].ctas++;
                    st[iter].pending += np;
                    st[iter].warps += (np + 31) / 32;
                }
                // threads of a CTA run concurrently: evaluate in descending order so that a thread mostly sees its
                // predecessor's OLD state (the adversarial interleaving)
                for (unsigned tt = 0; tt < cnt; tt++) {
                    const unsigned t = cnt - 1 - tt, i = i0 + t;
                    changed[t] = 0;
                    if (!pending[t]) continue;
                    EntCountSink sink;
                    const EntState st = ent_unpack(state[i - 1]);
                    const uint64_t v = ent_pack(ent_decode_range<false>(words, tabs, im.dcslot, im.acslot, im.dec_bpm,
                                                                       EntState{st.p, st.k, st.b, 0}, ent_sub_end(i, n, im.scan_bits), sink, &dummy));
                    changed[t] = ((v ^ mine[t]) & ENT_SYNC_MASK) != 0;
                    state[i] = v;
                    nvals[i] = sink.nvals;
                    mine[t] = v;
                    ever[t] |= changed[t];
                }
                bool more = false;
                for (unsigned t = 0; t < cnt; t++) {
                    pending[t] = t > 0 && changed[t - 1];
                    more |= pending[t];
                }
                if (!more) {
                    std::fill(changed.begin(), changed.end(), 0);
                    break;
                }
            }
            for (unsigned t = 0; t < cnt; t++) {
                const bool flag = changed[t] || (t == T - 1 && ever[t]);
                cout[i0 + t] = flag;
                any |= flag;
            }
        }
        std::swap(cin, cout);
        passes++;
        if (!any) break;
    }
    // prefix
    std::vector<uint32_t> first(n), first_val(n);
    uint32_t acc = 0, accv = 0;
    for (unsigned i = 0; i < n; i++) {
        first[i] = acc;
        first_val[i] = accv;
        acc += ent_unpack(state[i]).nb;
        accv += nvals[i];
    }
    // write
    bool completed = false;
    for (unsigned i = 0; i < n; i++) {
        if (first[i] >= im.total_blocks) continue;
        EntState st{0, 0, 0, 0};
        if (i) {
            st = ent_unpack(state[i - 1]);
            st.nb = 0;
        }
        EntCompactSink sink;
        sink.begin(r.cs.data(), &im, first[i], first_val[i], st.k == 0);
        const bool last = i + 1 == n;
        const uint32_t end = last ? im.scan_bits + ENT_TAIL_SLACK_BITS : ent_sub_end(i, n, im.scan_bits);
        unsigned bad = 0;
        const EntState e = ent_decode_range<true>(words, tabs, im.dcslot, im.acslot, im.dec_bpm, st, end, sink, &bad);
        sink.finish();
        if (sink.complete()) {
            completed = true;
            if (im.tight_end && e.p + 7 < im.scan_bits) bad |= ENT_BAD_TAIL;
        }
        else if (last) bad |= ENT_INCOMPLETE;
        else if (ent_pack(e) != state[i] || sink.vi - first_val[i] != nvals[i]) bad |= ENT_BAD_CHAIN;
        r.status |= bad;
    }
    if (!completed) r.status |= ENT_INCOMPLETE;
    // dc: differences -> values, per component in scan order, on the compact stream's dc array
    int16_t* dcarr = (int16_t*)(r.cs.data() + 8 * (size_t)im.nb_pad);
    for (int c = 0; c < d.ncomp; c++) {
        const unsigned hv = (unsigned)im.h[c] * im.v[c];
        uint16_t pred = 0;
        for (unsigned q = 0; q < im.comp_blocks[c]; q++) {
            int16_t* p = dcarr + (size_t)(im.mcu0 + q / hv) * im.bpm + im.comp_j0[c] + q % hv;
            pred = (uint16_t)(pred + (uint16_t)*p);
            *p = (int16_t)pred;
        }
    }
    r.passes = std::max(r.passes, passes);
}

struct HostResult {
    int status = 0;
    bool complete = false;
    std::vector<int16_t> coefs;
    b200jpg_image_desc desc;
};

static HostResult host_decode(const std::vector<uint8_t>& file) {
    HostResult r;
    memset(&r.desc, 0, sizeof r.desc);
    HostDecoder hd(file.data(), file.size());
    r.status = hd.entropy_decode();
    if (r.status != B200JPG_OK) return r;
    r.complete = true;
    for (size_t c = 0; c < hd.frame().comps.size(); c++) {
        if (!hd.component_has_data((int)c)) r.complete = false;
        else {
            const size_t n = (size_t)hd.frame().comps[c].block_w * hd.frame().comps[c].block_h * 64;
            r.coefs.insert(r.coefs.end(), hd.coefficients((int)c), hd.coefficients((int)c) + n);
        }
    }
    return r;
}

// 0 = fine, 1 = mismatch
static int check(const std::vector<uint8_t>& file, const char* name, int max_passes, bool verbose, unsigned* n_device, unsigned* n_flagged) {
    HostDecoder probe(file.data(), file.size());
    probe.probe_device_scan(true);
    const int st = probe.entropy_decode();
    if (st != B200JPG_INTERNAL_DEVICE_SCAN) {
        if (verbose) printf("%s: host path (status %d)\n", name, st);
        return 0;
    }
    b200jpg_image_desc d;
    memset(&d, 0, sizeof d);
    d.ncomp = (uint8_t)probe.frame().comps.size();
    for (int c = 0; c < d.ncomp; c++) d.comps[c] = probe.frame().comps[(size_t)c];
    const EmulResult e = emulate(file, probe, d, max_passes);
    if (!e.eligible) {
        if (verbose) printf("%s: host path (scan does not qualify)\n", name);
        return 0;
    }
    if (e.status) {
        (*n_flagged)++;
        if (verbose) printf("%s: flagged 0x%x after %u passes (%u subsequences) -> host\n", name, e.status, e.passes, e.nsub);
        return 0;
    }
    (*n_device)++;
    const HostResult h = host_decode(file);
    if (h.status != B200JPG_OK || !h.complete) {
        printf("%s: MISMATCH device accepted, host status %d complete %d\n", name, h.status, (int)h.complete);
        return 1;
    }
    if (h.coefs.size() != e.coefs.size() || memcmp(h.coefs.data(), e.coefs.data(), h.coefs.size() * 2) != 0) {
        size_t k = 0;
        while (k < h.coefs.size() && k < e.coefs.size() && h.coefs[k] == e.coefs[k]) k++;
        printf("%s: MISMATCH at coefficient %zu (block %zu pos %zu): host %d device %d\n", name, k, k / 64, k % 64, k < h.coefs.size() ? h.coefs[k] : -1,
               k < e.coefs.size() ? e.coefs[k] : -1);
        return 1;
    }
    if (verbose) printf("%s: device == host, %u subsequences in %u interval(s), %u sync passes\n", name, e.nsub, e.intervals, e.passes);
    return 0;
}

int main(int argc, char** argv) {
    int ncorrupt = 0, max_passes = 64;
    unsigned seed = 1;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--corrupt") && i + 2 < argc) {
            ncorrupt = atoi(argv[i + 1]);
            seed = (unsigned)atoi(argv[i + 2]);
            i += 2;
        } else if (!strcmp(argv[i], "--descending")) {
            g_cta_order = 1;
        } else if (!strcmp(argv[i], "--passes") && i + 1 < argc) {
            max_passes = atoi(argv[++i]);
        } else {
            files.push_back(argv[i]);
        }
    }
    int bad = 0;
    unsigned n_device = 0, n_flagged = 0;
    for (const auto& path : files) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) {
            printf("cannot open %s\n", path.c_str());
            return 2;
        }
        std::vector<uint8_t> data;
        uint8_t buf[65536];
        size_t got;
        while ((got = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + got);
        fclose(f);
        bad += check(data, path.c_str(), max_passes, true, &n_device, &n_flagged);
        if (ncorrupt > 0) {
            // locate the scan so that the edits land in entropy-coded data
            HostDecoder probe(data.data(), data.size());
            probe.probe_device_scan(true);
            if (probe.entropy_decode() != B200JPG_INTERNAL_DEVICE_SCAN) continue;
            const size_t begin = probe.device_scan().scan_begin;
            if (begin + 4 >= data.size()) continue;
            std::mt19937 rng(seed);
            for (int t = 0; t < ncorrupt; t++) {
                std::vector<uint8_t> c = data;
                const int kind = (int)(rng() % 4);
                const size_t at = begin + rng() % (c.size() - begin - 2);
                if (kind == 0) c[at] ^= (uint8_t)(1u << (rng() % 8));                      // one bit
                else if (kind == 1) c[at] = (uint8_t)rng();                                // one byte
                else if (kind == 2) c.erase(c.begin() + (long)at, c.begin() + (long)std::min(c.size() - 2, at + 1 + rng() % 64));  // drop a run
                else for (size_t q = at; q < std::min(c.size() - 2, at + 1 + rng() % 32); q++) c[q] = (uint8_t)rng();       // scramble a run
                char name[512];
                snprintf(name, sizeof name, "%s#%d", path.c_str(), t);
                bad += check(c, name, max_passes, false, &n_device, &n_flagged);
            }
        }
    }
    if (g_round_stats)
        for (size_t pass = 0; pass < g_rounds.size(); pass++)
            for (size_t it = 0; it < g_rounds[pass].size(); it++)
                printf("launch %zu round %zu: %u CTAs, %u subsequences re-decoded (%.1f per CTA), %u warps after gathering\n", pass + 1, it + 1,
                       g_rounds[pass][it].ctas, g_rounds[pass][it].pending, (double)g_rounds[pass][it].pending / g_rounds[pass][it].ctas, g_rounds[pass][it].warps);
#if defined(ENT_STATS)
    {
        unsigned long long tot = 0, acc = 0;
        for (int i = 0; i < 32; i++) tot += ent_stats_len[i];
        for (int i = 1; i <= 16; i++) {
            acc += ent_stats_len[i];
            printf("code length <= %2d: %.4f %%\n", i, 100.0 * (double)acc / (double)tot);
        }
    }
#endif
    printf("%s %u decoded on the emulated device, %u flagged for the host\n", bad ? "FAILED" : "ok", n_device, n_flagged);
    return bad ? 1 : 0;
}
