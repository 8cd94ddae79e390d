// host_decoder.h -- the host half of Decoder<R> (reference src/decoder.rs:101-1298): marker and
// segment parsing (src/parser.rs, src/marker.rs) and Huffman entropy decoding (src/huffman.rs)
// into dense per-component coefficient buffers -- the device input format.
// The reference is Rust; no Rust toolchain exists in this image, so the host side is C++.
//
// Differences from the reference's control flow (results are identical):
//  * coefficients of a baseline scan are written straight into the component's whole-image buffer
//    (block raster order) instead of one Vec per MCU row handed to Worker::append_row
//    (src/decoder.rs:1019-1060): the GPU worker wants one upload per component;
//  * a progressive component is snapshotted at the end of the scan that completes it
//    (src/decoder.rs:441-455, 1035-1048) -- the same coefficients the reference's worker sees.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/b200jpg.h"
#include "sbs.h"

namespace b200jpg {

struct HuffTable {  // src/huffman.rs:181-188
    bool present = false;
    uint8_t values[256];
    int nvalues = 0;
    int32_t delta[16];
    int32_t maxcode[16];
    uint8_t lut_value[256], lut_size[256];
    bool has_ac_lut = false;
    int16_t ac_value[256];
    uint8_t ac_run_size[256];
    bool build(const uint8_t bits[16], const uint8_t* vals, int nvals, bool is_ac);
};

struct FrameInfo {  // src/parser.rs:50-61
    bool is_baseline = false, is_differential = false, arithmetic = false;
    int coding_process = 0;  // B200JPG_CP_*
    uint8_t precision = 0;
    uint16_t image_w = 0, image_h = 0, output_w = 0, output_h = 0, mcu_w = 0, mcu_h = 0;
    std::vector<b200jpg_component> comps;
};

struct ScanInfo {  // src/parser.rs:64-74
    int n = 0;
    int comp_index[4], dc_table[4], ac_table[4];
    uint8_t ss_start = 0, ss_end = 0, ah = 0, al = 0;
};

// Filled at the first SOS when probing for device entropy decoding (entropy_dev.h): the scan is a complete
// sequential single-scan image (every component, in frame order, interleaved or a lone 1x1 component, all tables
// present) -- the case HostDecoder would write straight into a sparse stream.
struct DeviceScan {
    bool eligible = false;
    size_t scan_begin = 0;  // offset of the first entropy-coded byte in the file
    unsigned restart_interval = 0;  // MCUs between RSTn markers (DRI), 0 = none
    ScanInfo scan;
};
// positive (not an error): entropy_decode() stopped at the first SOS because the scan qualifies for the device
enum { B200JPG_INTERNAL_DEVICE_SCAN = 1 };

struct IccChunk {
    uint8_t num_markers, seq_no;
    std::vector<uint8_t> data;
};

class HostDecoder {
public:
    HostDecoder(const uint8_t* data, size_t len) : data_(data), len_(len) {}

    void set_color_transform(int ct) { has_ct_ = true; ct_ = ct; }
    void set_max_decoding_buffer_size(size_t m) { buffer_limit_ = m; }
    int read_info() { return decode_internal(true); }
    // all scans up to EOI; afterwards coefficients(i) are what the worker boundary receives
    int entropy_decode() { return decode_internal(false); }
    int scale(uint16_t req_w, uint16_t req_h, uint16_t* w, uint16_t* h);

    // Optional: entropy_decode() returns B200JPG_INTERNAL_DEVICE_SCAN at the first SOS of a qualifying scan instead of
    // decoding it; device_scan() then says where the entropy-coded bytes start.  Otherwise nothing changes.
    void probe_device_scan(bool on) { probe_device_ = on; }
    const DeviceScan& device_scan() const { return device_scan_; }
    const HuffTable& dc_table(int i) const { return dc_[i & 3]; }
    const HuffTable& ac_table(int i) const { return ac_[i & 3]; }
    // what decode_scan records for the worker when a component is finished by the (only) scan
    void capture_final_qtables();

    // nothing was set that the whole-file batch path does not know about (colour transform override, size limit, scaling)
    bool default_config() const {
        if (has_ct_ || buffer_limit_ != (size_t)-1) return false;
        for (const auto& c : frame_.comps)
            if (c.dct_scale != 8) return false;
        return true;
    }

    bool has_frame() const { return has_frame_; }
    const FrameInfo& frame() const { return frame_; }
    int determine_color_transform() const;  // src/decoder.rs:698-764
    int pixel_format() const;
    // component i was fed to the worker (finished, or rendered partially at EOI)
    bool component_has_data(int i) const { return i >= 0 && i < 4 && have_final_[i]; }
    const int16_t* coefficients(int i) const { return ext_[i] ? ext_[i] : final_[i].data(); }
    // Optional: caller-owned destination (e.g. page-locked memory) for component i's final coefficients,
    // block_w*block_h*64 int16; must be set after read_info() and before entropy_decode().
    void set_external_buffer(int i, int16_t* p) { if (i >= 0 && i < 4) ext_[i] = p; }
    // Optional: sparse block stream destination (sbs.h) of at least SbsLayout::make(total blocks).worst_bytes()
    // bytes; must be set after read_info() and before entropy_decode().  A single-scan sequential image is
    // written straight from the Huffman loop in scan order; anything else (progressive, several scans) is decoded
    // densely as usual and compacted at the end, component by component.
    void set_sbs_sink(uint8_t* base) { sbs_base_ = base; }
    size_t sbs_length() const { return sbs_len_; }    // valid after a successful entropy_decode()
    unsigned sbs_order() const { return sbs_direct_ ? SBS_INTERLEAVED : SBS_PLANAR; }
    size_t total_blocks() const {
        size_t n = 0;
        for (const auto& c : frame_.comps) n += (size_t)c.block_w * c.block_h;
        return n;
    }
    // quantisation table captured when the component was handed to the worker (RowData, src/decoder.rs:850-857)
    const uint16_t* component_qtable(int i) const { return final_qt_[i]; }
    bool buffer_limit_exceeded() const;
    bool icc_profile(std::vector<uint8_t>* out) const;  // src/decoder.rs:211-241
    const std::vector<uint8_t>* exif() const { return has_exif_ ? &exif_ : nullptr; }
    const std::vector<uint8_t>* xmp() const { return has_xmp_ ? &xmp_ : nullptr; }
    const std::string& error() const { return err_; }

private:
    int decode_internal(bool stop_after_metadata);
    int fail(int code, const char* fmt, ...);
    // reader
    int read_u8(uint8_t* b);
    int read_u16(uint16_t* v);
    int read_exact(uint8_t* dst, size_t n);
    int skip(size_t n);
    int read_length(size_t* len);
    int read_marker(uint8_t* m);
    // segments
    int parse_sof(uint8_t marker);
    int parse_sos(ScanInfo* s);
    int parse_dqt();
    int parse_dht();
    int parse_app(uint8_t marker);
    void fill_default_mjpeg_tables(const ScanInfo& s);
    // entropy decoding
    int decode_scan(const ScanInfo& scan, const bool finished[4], bool* has_marker, uint8_t* marker);
    int read_bits();
    int get_bits(uint8_t count, uint16_t* v);
    int receive_extend(uint8_t count, int16_t* v);
    int huff_decode(const HuffTable& t, uint8_t* out);
    int take_marker(bool* has, uint8_t* m);
    int decode_block(int16_t* c, uint64_t* nz, const HuffTable& dc, const HuffTable& ac, const ScanInfo& s, uint16_t* eob_run, int16_t* pred);
    template <class Sink>
    int decode_block_seq(Sink& sink, const HuffTable& dc, const HuffTable& ac, uint16_t* eob_run, int16_t* pred);
    int finish_sbs();
    int decode_block_sa(int16_t* c, uint64_t* nz, const HuffTable& ac, const ScanInfo& s, uint16_t* eob_run);
    int refine_non_zeroes(int16_t* c, uint8_t start, uint8_t end, uint8_t zrl, int16_t bit, uint8_t* ret);
    int refine_non_zeroes_map(int16_t* c, uint64_t nz, uint8_t start, uint8_t end, uint8_t zrl, int16_t bit, uint8_t* ret);

    const uint8_t* data_;
    size_t len_, pos_ = 0;
    bool has_frame_ = false;
    FrameInfo frame_;
    HuffTable dc_[4], ac_[4];
    bool has_qt_[4] = {false, false, false, false};
    uint16_t qt_[4][64];
    uint16_t restart_interval_ = 0;
    bool has_adobe_ = false;
    int adobe_ = 0;
    bool has_ct_ = false;
    int ct_ = 0;
    bool is_jfif_ = false, is_mjpeg_ = false;
    std::vector<IccChunk> icc_;
    std::vector<uint8_t> exif_, xmp_;
    bool has_exif_ = false, has_xmp_ = false;
    size_t buffer_limit_ = (size_t)-1;
    // progressive working store (src/decoder.rs:124-126) and what the worker gets
    std::vector<int16_t> work_[4];
    // progressive frames: per block of work_, bit k = the coefficient with zig-zag index k is non-zero.  The refinement scans
    // (src/decoder.rs:1174-1298) walk "the non-zero coefficients of the band" and "the n-th zero coefficient" for every block
    // of every scan; with the map both are bit scans over one word instead of 63 scattered loads per block and scan
    std::vector<uint64_t> nz_[4];
    bool has_work_ = false;
    uint64_t finished_mask_[4] = {0, 0, 0, 0};
    std::vector<int16_t> final_[4];
    int16_t* ext_[4] = {nullptr, nullptr, nullptr, nullptr};
    bool have_final_[4] = {false, false, false, false};
    uint16_t final_qt_[4][64];
    // sparse block stream output
    uint8_t* sbs_base_ = nullptr;
    SbsWriter sbs_;
    bool sbs_direct_ = false;
    size_t sbs_len_ = 0;
    bool probe_device_ = false;
    DeviceScan device_scan_;
    bool device_scan_ok(const ScanInfo& scan) const;
    // bit reader (src/huffman.rs:14-18)
    uint64_t bits_ = 0;
    uint8_t num_bits_ = 0;
    bool has_marker_ = false;
    uint8_t marker_ = 0;
    std::string err_;
};

}  // namespace b200jpg
