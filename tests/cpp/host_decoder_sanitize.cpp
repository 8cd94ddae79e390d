// host_decoder_sanitize.cpp -- the host half of the decoder (csrc/host_decoder.cpp: marker parser, Huffman decoder,
// scan driver, sparse-stream writer, entropy payload builder) over a list of files, meant to be built with
// -fsanitize=address,undefined (tests/test_host_sanitize.py).  The analogue of the reference's crash tests
// (tests/crashtest/mod.rs: "decoding must not panic") with the sanitizers standing in for Rust's bounds checks.
// Every file is run through: read_info, a dense entropy decode, a sparse-stream entropy decode, the device-scan probe +
// payload builder, and the three scaled sizes.  Exit code 0 = no sanitizer report (the process aborts otherwise).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../jpeg_decoder_b200/csrc/entropy_host.h"
#include "../../jpeg_decoder_b200/csrc/host_decoder.h"
#include "../../jpeg_decoder_b200/csrc/sbs.h"

using namespace b200jpg;

static std::vector<uint8_t> read_file(const char* path) {
    std::vector<uint8_t> v;
    FILE* f = fopen(path, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize((size_t)(n > 0 ? n : 0));
    if (n > 0 && fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    unsigned ok = 0, failed = 0, probed = 0;
    for (int i = 1; i < argc; i++) {
        const std::vector<uint8_t> file = read_file(argv[i]);
        {
            HostDecoder d(file.data(), file.size());
            d.read_info();
        }
        {
            HostDecoder d(file.data(), file.size());
            d.set_max_decoding_buffer_size((size_t)1 << 25);  // crafted headers may announce 65535 x 65535
            const int rc = d.entropy_decode();
            rc == B200JPG_OK ? ok++ : failed++;
            if (rc == B200JPG_OK) {
                d.determine_color_transform();
                std::vector<uint8_t> icc;
                d.icc_profile(&icc);
            }
        }
        {
            HostDecoder d(file.data(), file.size());
            d.set_max_decoding_buffer_size((size_t)1 << 25);
            if (d.read_info() == B200JPG_OK && d.has_frame() && d.total_blocks() < (1u << 22)) {
                std::vector<uint8_t> sbs(SbsLayout::make(d.total_blocks()).worst_bytes() + 64);
                d.set_sbs_sink(sbs.data());
                d.entropy_decode();
            }
        }
        {
            HostDecoder d(file.data(), file.size());
            d.probe_device_scan(true);
            if (d.entropy_decode() == B200JPG_INTERNAL_DEVICE_SCAN) {
                probed++;
                std::vector<uint8_t> payload(ent_payload_bound(file.size()) + 64);
                ent_build_payload(d, file.data(), file.size(), payload.data(), payload.size());
            }
        }
        for (unsigned s = 1; s <= 4; s *= 2) {
            HostDecoder d(file.data(), file.size());
            d.set_max_decoding_buffer_size((size_t)1 << 25);
            if (d.read_info() != B200JPG_OK || !d.has_frame()) break;
            uint16_t w = 0, h = 0;
            d.scale((uint16_t)((d.frame().image_w * s + 7) / 8), (uint16_t)((d.frame().image_h * s + 7) / 8), &w, &h);
            d.entropy_decode();
        }
    }
    printf("ok %u decoded, %u rejected, %u device-scan payloads built, %d files\n", ok, failed, probed, argc - 1);
    return 0;
}
