// wav.hpp — N4 output sink: RIFF/WAVE writer for rendered blocks (IEEE float32, WAVE_FORMAT_IEEE_FLOAT).
// The reference scopes file output out of the library (README.md:22-26: the client owns it) and hands buffers to
// Client::audio_rendered (src/client/client.rs:8-15) as planar [n_slots x n_times]; a WAV file wants interleaved frames.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace frb::host {

class WavWriter {
public:
    // header sizes are patched at close(); files are limited to the 4 GiB of plain RIFF
    bool open(const std::string& path, uint32_t n_channels, uint32_t sample_rate, std::string* err) {
        if (n_channels == 0 || n_channels > 65535 || sample_rate == 0) { if (err) *err = "wav: bad channel count / sample rate"; return false; }
        f_ = std::fopen(path.c_str(), "wb");
        if (!f_) { if (err) *err = "wav: cannot create " + path; return false; }
        ch_ = n_channels; sr_ = sample_rate; frames_ = 0;
        return write_header();
    }
    // block: planar [n_slots x n_times] (slot s at block[s * n_times + i]); n_slots must equal the channel count
    bool write(const float* block, uint32_t n_slots, uint64_t n_times, std::string* err) {
        if (!f_) { if (err) *err = "wav: not open"; return false; }
        if (n_slots != ch_) { if (err) *err = "wav: block has " + std::to_string(n_slots) + " slots, file has " + std::to_string(ch_) + " channels"; return false; }
        if ((frames_ + n_times) * ch_ * 4 + 58 > 0xffffffffull) { if (err) *err = "wav: file would exceed 4 GiB"; return false; }
        if (ch_ == 1) {
            if (std::fwrite(block, 4, n_times, f_) != n_times) { if (err) *err = "wav: write failed"; return false; }
        } else {
            constexpr uint64_t TILE = 4096;
            buf_.resize(TILE * ch_);
            for (uint64_t i0 = 0; i0 < n_times; i0 += TILE) {
                const uint64_t m = std::min(TILE, n_times - i0);
                for (uint32_t c = 0; c < ch_; c++) {
                    const float* src = block + (uint64_t)c * n_times + i0;
                    for (uint64_t i = 0; i < m; i++) buf_[i * ch_ + c] = src[i];
                }
                if (std::fwrite(buf_.data(), 4, m * ch_, f_) != m * ch_) { if (err) *err = "wav: write failed"; return false; }
            }
        }
        frames_ += n_times;
        return true;
    }
    bool close() {
        if (!f_) return true;
        bool ok = std::fseek(f_, 0, SEEK_SET) == 0 && write_header();
        ok = (std::fclose(f_) == 0) && ok;
        f_ = nullptr;
        return ok;
    }
    uint64_t frames() const { return frames_; }
    ~WavWriter() { close(); }

private:
    static void put32(std::vector<uint8_t>& h, uint32_t v) { for (int i = 0; i < 4; i++) h.push_back((uint8_t)(v >> (8 * i))); }
    static void put16(std::vector<uint8_t>& h, uint16_t v) { h.push_back((uint8_t)v); h.push_back((uint8_t)(v >> 8)); }
    static void tag(std::vector<uint8_t>& h, const char* t) { for (int i = 0; i < 4; i++) h.push_back((uint8_t)t[i]); }
    bool write_header() {
        const uint32_t data_bytes = (uint32_t)(frames_ * ch_ * 4);
        std::vector<uint8_t> h;
        tag(h, "RIFF"); put32(h, 4 + (8 + 18) + (8 + 4) + (8 + data_bytes)); tag(h, "WAVE");
        tag(h, "fmt "); put32(h, 18);
        put16(h, 3);                                   // WAVE_FORMAT_IEEE_FLOAT
        put16(h, (uint16_t)ch_); put32(h, sr_); put32(h, sr_ * ch_ * 4); put16(h, (uint16_t)(ch_ * 4)); put16(h, 32);
        put16(h, 0);                                   // cbSize (non-PCM formats carry it)
        tag(h, "fact"); put32(h, 4); put32(h, (uint32_t)frames_);
        tag(h, "data"); put32(h, data_bytes);
        return std::fwrite(h.data(), 1, h.size(), f_) == h.size();   // 58 bytes
    }
    FILE* f_ = nullptr;
    uint32_t ch_ = 0, sr_ = 0;
    uint64_t frames_ = 0;
    std::vector<float> buf_;
};

}  // namespace frb::host
