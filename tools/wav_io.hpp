// wav_io.hpp — RIFF/WAVE reader (PCM 8/16/24/32-bit, IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE) and float32 writer.
// The reference's offline job reads its voice file frame by frame as ONE float (WAVReader::ReadAsOneFloat,
// source/rt.resynth.lib.params.cpp:297-314; the reader itself lives in the absent cpp.algorithms checkout — SPEC S11:
// one float per frame = the mean of the frame's channels) and writes IEEE-float stereo (:342-356).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace wav_io {

struct Audio {
  int sample_rate = 0;
  int channels = 0;
  std::vector<float> samples;  // interleaved
  size_t frames() const { return channels ? samples.size() / channels : 0; }
};

inline uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }
inline uint16_t rd16(const unsigned char* p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }

inline Audio read(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open '" + path + "'");
  std::vector<unsigned char> buf;
  {
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? static_cast<size_t>(n) : 0);
    if (n > 0 && std::fread(buf.data(), 1, buf.size(), f) != buf.size()) {
      std::fclose(f);
      throw std::runtime_error("short read on '" + path + "'");
    }
    std::fclose(f);
  }
  if (buf.size() < 12 || std::memcmp(buf.data(), "RIFF", 4) || std::memcmp(buf.data() + 8, "WAVE", 4))
    throw std::runtime_error("'" + path + "' is not a RIFF/WAVE file");
  int format = 0, bits = 0;
  Audio a;
  const unsigned char* data = nullptr;
  size_t data_bytes = 0;
  size_t pos = 12;
  while (pos + 8 <= buf.size()) {
    const unsigned char* ck = buf.data() + pos;
    const uint32_t len = rd32(ck + 4);
    const size_t body = pos + 8;
    if (!std::memcmp(ck, "fmt ", 4)) {
      if (len < 16 || body + 16 > buf.size()) throw std::runtime_error("bad fmt chunk");
      format = rd16(buf.data() + body);
      a.channels = rd16(buf.data() + body + 2);
      a.sample_rate = static_cast<int>(rd32(buf.data() + body + 4));
      bits = rd16(buf.data() + body + 14);
      if (format == 0xFFFE && len >= 26) format = rd16(buf.data() + body + 24);  // extensible: first word of the sub-format GUID
    } else if (!std::memcmp(ck, "data", 4)) {
      data = buf.data() + body;
      data_bytes = std::min<size_t>(len, buf.size() - body);
    }
    pos = body + len + (len & 1);
  }
  if (!data || !a.channels || !bits) throw std::runtime_error("'" + path + "': missing fmt or data chunk");
  const bool pcm_ok = format == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32);
  const bool float_ok = format == 3 && (bits == 32 || bits == 64);
  if (!pcm_ok && !float_ok) throw std::runtime_error("'" + path + "': unsupported sample format");
  const int bytes = bits / 8;
  const size_t n = data_bytes / bytes;
  a.samples.resize(n - n % a.channels);
  for (size_t i = 0; i < a.samples.size(); ++i) {
    const unsigned char* p = data + i * bytes;
    float v = 0.f;
    if (format == 3 && bits == 32) {
      std::memcpy(&v, p, 4);
    } else if (format == 3 && bits == 64) {
      double d;
      std::memcpy(&d, p, 8);
      v = static_cast<float>(d);
    } else if (format == 1 && bits == 8) {
      v = (static_cast<int>(p[0]) - 128) / 128.f;
    } else if (format == 1 && bits == 16) {
      v = static_cast<int16_t>(rd16(p)) / 32768.f;
    } else if (format == 1 && bits == 24) {
      int32_t x = p[0] | (p[1] << 8) | (p[2] << 16);
      if (x & 0x800000) x |= ~0xFFFFFF;
      v = x / 8388608.f;
    } else if (format == 1 && bits == 32) {
      v = static_cast<int32_t>(rd32(p)) / 2147483648.f;
    } else {
      throw std::runtime_error("'" + path + "': unsupported sample format");
    }
    a.samples[i] = v;
  }
  return a;
}

// one float per frame: the mean of the frame's channels (SPEC S11)
inline std::vector<float> as_one_float(const Audio& a) {
  std::vector<float> mono(a.frames());
  for (size_t i = 0; i < mono.size(); ++i) {
    double s = 0;
    for (int c = 0; c < a.channels; ++c) s += a.samples[i * a.channels + c];
    mono[i] = static_cast<float>(s / a.channels);
  }
  return mono;
}

inline void write_float32(const std::string& path, int sample_rate, int channels, const float* interleaved, size_t frames) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot create '" + path + "'");
  if (frames * static_cast<size_t>(channels) * 4 > 0xFFFFFFFFull - 36)
    throw std::runtime_error("'" + path + "': more than 4 GiB of samples do not fit a RIFF/WAVE header");
  const uint32_t data_bytes = static_cast<uint32_t>(frames * channels * 4);
  unsigned char h[44];
  auto w32 = [&](int o, uint32_t v) { h[o] = v & 255; h[o + 1] = (v >> 8) & 255; h[o + 2] = (v >> 16) & 255; h[o + 3] = (v >> 24) & 255; };
  auto w16 = [&](int o, uint16_t v) { h[o] = v & 255; h[o + 1] = (v >> 8) & 255; };
  std::memcpy(h, "RIFF", 4);
  w32(4, 36 + data_bytes);
  std::memcpy(h + 8, "WAVEfmt ", 8);
  w32(16, 16);
  w16(20, 3);  // IEEE float
  w16(22, static_cast<uint16_t>(channels));
  w32(24, static_cast<uint32_t>(sample_rate));
  w32(28, static_cast<uint32_t>(sample_rate * channels * 4));
  w16(32, static_cast<uint16_t>(channels * 4));
  w16(34, 32);
  std::memcpy(h + 36, "data", 4);
  w32(40, data_bytes);
  const bool ok = std::fwrite(h, 1, 44, f) == 44 && std::fwrite(interleaved, 4, frames * channels, f) == frames * channels;
  std::fclose(f);
  if (!ok) throw std::runtime_error("short write on '" + path + "'");
}

}  // namespace wav_io
