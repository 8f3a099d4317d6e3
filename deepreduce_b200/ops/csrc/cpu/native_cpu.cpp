// Host-side native ops (no CUDA dependency) — the equivalents of the
// reference's C++ TensorFlow CPU ops and their un-vendored third-party
// libraries:
//   * bloom filter + selection policies leftmost / random / p0 / conflict-sets
//     (reference tensorflow/bloom_filter_compression.cc, policies.hpp,
//      third_party/bloomfilter OrdinaryBloomFilter)
//   * integer-array codecs chosen by id (reference integer_compression.cc over
//     third_party/FastPFor): copy, vbyte, bp32, bp128, simple8b, pfor128
//   * CSV value/coefficient logger (reference logger.cc, compression_utils.hpp)
// Hashing / bit layout follow deepreduce_b200/spec.py exactly.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

namespace py = pybind11;

namespace {

constexpr uint32_t kGolden = 0x9E3779B1u, kBAdd = 0x7F4A7C15u;

inline uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
inline void hash_ab(uint32_t x, uint32_t seed, uint32_t& a, uint32_t& b) {
  uint32_t y = x ^ seed;
  a = fmix32(y);
  b = fmix32(y * kGolden + kBAdd) | 1u;
}
inline uint32_t mulhi(uint32_t h, uint32_t m) { return (uint32_t)(((uint64_t)h * m) >> 32); }
inline uint32_t policy_hash(uint32_t x, uint32_t seed) { return fmix32((x * kGolden + seed) ^ 0x5BD1E995u); }

struct Bloom {
  std::vector<uint32_t> words;
  uint32_t k, m_bits, seed;
  Bloom(uint32_t k_, uint32_t m_, uint32_t s_) : words((m_ + 31) / 32, 0u), k(k_), m_bits(m_), seed(s_) {}
  void insert(uint32_t x) {
    uint32_t a, b; hash_ab(x, seed, a, b);
    for (uint32_t j = 0; j < k; ++j) { uint32_t p = mulhi(a + j * b, m_bits); words[p >> 5] |= 1u << (p & 31); }
  }
  bool query(uint32_t x) const {
    uint32_t a, b; hash_ab(x, seed, a, b);
    for (uint32_t j = 0; j < k; ++j) { uint32_t p = mulhi(a + j * b, m_bits); if (!((words[p >> 5] >> (p & 31)) & 1u)) return false; }
    return true;
  }
  uint32_t pos(uint32_t x, uint32_t j) const { uint32_t a, b; hash_ab(x, seed, a, b); return mulhi(a + j * b, m_bits); }
};

py::array_t<uint32_t> bloom_insert(py::array_t<int64_t, py::array::c_style | py::array::forcecast> idx, uint32_t k,
                                   uint32_t m_bits, uint32_t seed) {
  Bloom bf(k, m_bits, seed);
  auto r = idx.unchecked<1>();
  for (py::ssize_t i = 0; i < r.shape(0); ++i) bf.insert((uint32_t)r(i));
  py::array_t<uint32_t> out(bf.words.size());
  std::memcpy(out.mutable_data(), bf.words.data(), bf.words.size() * 4);
  return out;
}

// P2 — conflict sets (paper Alg. 1; reference policies.hpp:43-146).  Tie-breaks
// are specified in codecs/bloom.py::conflict_sets_oracle and mirrored here.
std::vector<int64_t> conflict_sets_impl(const std::vector<int64_t>& P, int64_t K, const Bloom& bf, uint32_t pseed) {
  std::map<uint32_t, std::vector<int64_t>> sets;
  for (int64_t x : P)
    for (uint32_t j = 0; j < bf.k; ++j) {
      auto& s = sets[bf.pos((uint32_t)x, j)];
      if (s.empty() || s.back() != x) s.push_back(x);
    }
  std::vector<std::pair<uint32_t, std::vector<int64_t>>> ord(sets.begin(), sets.end());
  std::stable_sort(ord.begin(), ord.end(), [](const auto& l, const auto& r) {
    return l.second.size() != r.second.size() ? l.second.size() < r.second.size() : l.first < r.first;
  });
  std::unordered_set<int64_t> chosen;
  int64_t left = std::min<int64_t>(K, (int64_t)P.size());
  uint32_t draw = 0;
  while (left > 0) {
    bool picked = false;
    for (auto& kv : ord) {
      if (left == 0) break;
      auto& cs = kv.second;
      const size_t before = cs.size();
      cs.erase(std::remove_if(cs.begin(), cs.end(), [&](int64_t x) { return chosen.count(x) != 0; }), cs.end());
      if (cs.size() == before && !cs.empty()) {
        const uint32_t r = policy_hash(draw++, pseed) % (uint32_t)cs.size();
        chosen.insert(cs[r]);
        cs.erase(cs.begin() + r);
        --left;
        picked = true;
      }
    }
    if (!picked) {
      for (int64_t x : P) {
        if (left == 0) break;
        if (!chosen.count(x)) { chosen.insert(x); --left; }
      }
    }
  }
  std::vector<int64_t> out(chosen.begin(), chosen.end());
  std::sort(out.begin(), out.end());
  return out;
}

// policy: 0 leftmost, 1 random (seeded hash rank), 2 p0, 3 conflict sets
py::array_t<int64_t> bloom_select(py::array_t<uint32_t, py::array::c_style | py::array::forcecast> words, int64_t d,
                                  int64_t K, uint32_t k, uint32_t m_bits, uint32_t seed, int policy, uint32_t pseed) {
  Bloom bf(k, m_bits, seed);
  if ((size_t)words.size() < bf.words.size()) throw std::runtime_error("filter too short");
  std::memcpy(bf.words.data(), words.data(), bf.words.size() * 4);
  std::vector<int64_t> sel;
  if (policy == 0) {
    for (int64_t i = 0; i < d && (int64_t)sel.size() < K; ++i) if (bf.query((uint32_t)i)) sel.push_back(i);
  } else {
    std::vector<int64_t> P;
    for (int64_t i = 0; i < d; ++i) if (bf.query((uint32_t)i)) P.push_back(i);
    if (policy == 2 || ((int64_t)P.size() <= K && policy != 3)) {
      sel.swap(P);
    } else if (policy == 1) {
      std::vector<uint64_t> keyed(P.size());
      for (size_t i = 0; i < P.size(); ++i) keyed[i] = ((uint64_t)policy_hash((uint32_t)P[i], pseed) << 32) | (uint64_t)P[i];
      std::nth_element(keyed.begin(), keyed.begin() + K, keyed.end());
      sel.resize(K);
      for (int64_t i = 0; i < K; ++i) sel[i] = (int64_t)(keyed[i] & 0xFFFFFFFFull);
      std::sort(sel.begin(), sel.end());
    } else {
      sel = conflict_sets_impl(P, K, bf, pseed);
    }
  }
  py::array_t<int64_t> out(sel.size());
  if (!sel.empty()) std::memcpy(out.mutable_data(), sel.data(), sel.size() * 8);
  return out;
}

py::array_t<int64_t> conflict_sets(py::array_t<int64_t, py::array::c_style | py::array::forcecast> positives, int64_t K,
                                   uint32_t k, uint32_t m_bits, uint32_t seed, uint32_t pseed) {
  Bloom bf(k, m_bits, seed);
  std::vector<int64_t> P(positives.data(), positives.data() + positives.size());
  auto sel = conflict_sets_impl(P, K, bf, pseed);
  py::array_t<int64_t> out(sel.size());
  if (!sel.empty()) std::memcpy(out.mutable_data(), sel.data(), sel.size() * 8);
  return out;
}

// ---------------------------------------------------------------------------
// integer codecs
// ---------------------------------------------------------------------------
inline uint32_t bitwidth(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

void pack_block(const uint32_t* in, uint32_t n, uint32_t width, std::vector<uint32_t>& out) {
  const size_t base = out.size();
  const uint32_t nw = (n * width + 31) / 32;
  out.resize(base + nw, 0u);
  uint64_t bit = 0;
  for (uint32_t i = 0; i < n; ++i, bit += width) {
    if (!width) continue;
    const uint64_t v = in[i];
    const uint32_t w = (uint32_t)(bit >> 5), s = (uint32_t)(bit & 31);
    out[base + w] |= (uint32_t)(v << s);
    if (s + width > 32) out[base + w + 1] |= (uint32_t)(v >> (32 - s));
  }
}

void unpack_block(const uint32_t* in, uint32_t n, uint32_t width, uint32_t* out) {
  uint64_t bit = 0;
  const uint64_t mask = width >= 32 ? 0xFFFFFFFFull : ((1ull << width) - 1);
  const uint32_t nw = (n * width + 31) / 32;
  for (uint32_t i = 0; i < n; ++i, bit += width) {
    if (!width) { out[i] = 0; continue; }
    const uint32_t w = (uint32_t)(bit >> 5), s = (uint32_t)(bit & 31);
    uint64_t win = in[w];
    if (w + 1 < nw) win |= (uint64_t)in[w + 1] << 32;
    out[i] = (uint32_t)((win >> s) & mask);
  }
}

std::vector<uint32_t> enc_bp(const uint32_t* a, size_t n, uint32_t block) {
  std::vector<uint32_t> out;
  std::vector<uint32_t> tmp(block);
  for (size_t lo = 0; lo < n; lo += block) {
    const uint32_t cnt = (uint32_t)std::min<size_t>(block, n - lo);
    std::fill(tmp.begin(), tmp.end(), 0u);
    std::memcpy(tmp.data(), a + lo, cnt * 4);
    uint32_t m = 0;
    for (uint32_t i = 0; i < block; ++i) m |= tmp[i];
    const uint32_t width = bitwidth(m);
    out.push_back(width);
    pack_block(tmp.data(), block, width, out);
  }
  return out;
}

std::vector<uint32_t> dec_bp(const uint32_t* w, size_t nw, size_t n, uint32_t block) {
  std::vector<uint32_t> out(((n + block - 1) / block) * block);
  size_t p = 0;
  for (size_t lo = 0; lo < n; lo += block) {
    if (p >= nw) throw std::runtime_error("bp: truncated input");
    const uint32_t width = w[p++];
    unpack_block(w + p, block, width, out.data() + lo);
    p += (block * width + 31) / 32;
  }
  out.resize(n);
  return out;
}

std::vector<uint32_t> enc_vbyte(const uint32_t* a, size_t n) {
  std::vector<uint8_t> b;
  b.reserve(n * 2);
  for (size_t i = 0; i < n; ++i) {
    uint32_t v = a[i];
    while (v >= 128) { b.push_back((uint8_t)((v & 127) | 128)); v >>= 7; }
    b.push_back((uint8_t)v);
  }
  while (b.size() % 4) b.push_back(0);
  std::vector<uint32_t> out(b.size() / 4);
  std::memcpy(out.data(), b.data(), b.size());
  return out;
}

std::vector<uint32_t> dec_vbyte(const uint32_t* w, size_t nw, size_t n) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(w);
  const size_t nb = nw * 4;
  std::vector<uint32_t> out(n);
  size_t p = 0;
  for (size_t i = 0; i < n; ++i) {
    uint32_t v = 0, s = 0;
    while (true) {
      if (p >= nb) throw std::runtime_error("vbyte: truncated input");
      const uint8_t c = b[p++];
      v |= (uint32_t)(c & 127) << s;
      s += 7;
      if (c < 128) break;
    }
    out[i] = v;
  }
  return out;
}

// simple8b: 64-bit words, 4-bit selector + 60 data bits; values must fit in 60 bits (always true for u32)
const uint32_t kS8bN[16] = {240, 120, 60, 30, 20, 15, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1};
const uint32_t kS8bB[16] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60};

std::vector<uint32_t> enc_simple8b(const uint32_t* a, size_t n) {
  std::vector<uint64_t> out;
  size_t i = 0;
  while (i < n) {
    int sel = 15;
    for (int s = 0; s < 16; ++s) {
      const uint32_t cnt = kS8bN[s], bits = kS8bB[s];
      if (bits == 0 && i + cnt > n) continue;         // run selectors need a full run
      const size_t take = std::min<size_t>(cnt, n - i);
      bool ok = true;
      for (size_t j = 0; j < take && ok; ++j) {
        if (bits == 0) ok = (a[i + j] == 1);          // selectors 0,1: runs of the value 1 (dense gap streams)
        else if (bits < 32) ok = a[i + j] < (1u << bits);
      }
      if (ok) { sel = s; break; }
    }
    const uint32_t cnt = kS8bN[sel], bits = kS8bB[sel];
    const size_t take = std::min<size_t>(cnt, n - i);
    uint64_t w = (uint64_t)sel << 60;
    for (size_t j = 0; j < take && bits; ++j) w |= (uint64_t)a[i + j] << (j * bits);
    out.push_back(w);
    i += take;
  }
  std::vector<uint32_t> o32(out.size() * 2);
  std::memcpy(o32.data(), out.data(), out.size() * 8);
  return o32;
}

std::vector<uint32_t> dec_simple8b(const uint32_t* w32, size_t nw, size_t n) {
  std::vector<uint32_t> out(n);
  size_t i = 0, p = 0;
  while (i < n) {
    if (p + 2 > nw) throw std::runtime_error("simple8b: truncated input");
    uint64_t w; std::memcpy(&w, w32 + p, 8); p += 2;
    const int sel = (int)(w >> 60);
    const uint32_t cnt = kS8bN[sel], bits = kS8bB[sel];
    const uint64_t mask = bits ? ((bits >= 64 ? ~0ull : (1ull << bits) - 1)) : 0;
    for (uint32_t j = 0; j < cnt && i < n; ++j, ++i) out[i] = bits ? (uint32_t)((w >> (j * bits)) & mask) : 1u;
  }
  return out;
}

// pfor128: per block of 128 pick width b minimising b*128 + exceptions*(8+32);
// low b bits are packed, exceptions stored as (position:u8 ..., high bits:u32 ...).
std::vector<uint32_t> enc_pfor(const uint32_t* a, size_t n) {
  const uint32_t B = 128;
  std::vector<uint32_t> out, tmp(B), low(B);
  for (size_t lo = 0; lo < n; lo += B) {
    const uint32_t cnt = (uint32_t)std::min<size_t>(B, n - lo);
    std::fill(tmp.begin(), tmp.end(), 0u);
    std::memcpy(tmp.data(), a + lo, cnt * 4);
    uint32_t hist[33] = {0};
    for (uint32_t i = 0; i < B; ++i) hist[bitwidth(tmp[i])]++;
    uint32_t best_b = 32, best_cost = 32 * B, exc = 0;
    for (int b = 32; b >= 0; --b) {
      const uint32_t cost = b * B + exc * 40;
      if (cost < best_cost) { best_cost = cost; best_b = b; }
      exc += hist[b];
    }
    std::vector<uint8_t> pos;
    std::vector<uint32_t> high;
    for (uint32_t i = 0; i < B; ++i) {
      if (bitwidth(tmp[i]) > best_b) { pos.push_back((uint8_t)i); high.push_back(best_b >= 32 ? 0 : tmp[i] >> best_b); }
      low[i] = best_b >= 32 ? tmp[i] : (tmp[i] & ((1u << best_b) - 1));
    }
    out.push_back(best_b | ((uint32_t)pos.size() << 8));
    pack_block(low.data(), B, best_b, out);
    while (pos.size() % 4) pos.push_back(0);
    const size_t base = out.size();
    out.resize(base + pos.size() / 4);
    std::memcpy(out.data() + base, pos.data(), pos.size());
    out.insert(out.end(), high.begin(), high.end());
  }
  return out;
}

std::vector<uint32_t> dec_pfor(const uint32_t* w, size_t nw, size_t n) {
  const uint32_t B = 128;
  std::vector<uint32_t> out(((n + B - 1) / B) * B);
  size_t p = 0;
  for (size_t lo = 0; lo < n; lo += B) {
    if (p >= nw) throw std::runtime_error("pfor: truncated input");
    const uint32_t head = w[p++];
    const uint32_t b = head & 0xFF, ne = head >> 8;
    unpack_block(w + p, B, b, out.data() + lo);
    p += (B * b + 31) / 32;
    const uint8_t* pos = reinterpret_cast<const uint8_t*>(w + p);
    p += (ne + 3) / 4;
    for (uint32_t e = 0; e < ne; ++e) out[lo + pos[e]] |= (b >= 32 ? 0u : w[p + e] << b);
    p += ne;
  }
  out.resize(n);
  return out;
}

py::array_t<uint32_t> to_np(const std::vector<uint32_t>& v) {
  py::array_t<uint32_t> out(v.size());
  if (!v.empty()) std::memcpy(out.mutable_data(), v.data(), v.size() * 4);
  return out;
}

py::array_t<uint32_t> int_encode(int codec, py::array_t<uint32_t, py::array::c_style | py::array::forcecast> a) {
  const uint32_t* p = a.data();
  const size_t n = (size_t)a.size();
  switch (codec) {
    case 0: return to_np(std::vector<uint32_t>(p, p + n));
    case 1: return to_np(enc_vbyte(p, n));
    case 2: return to_np(enc_bp(p, n, 32));
    case 3: return to_np(enc_bp(p, n, 128));
    case 4: return to_np(enc_simple8b(p, n));
    case 5: return to_np(enc_pfor(p, n));
  }
  throw std::runtime_error("unknown integer codec id");
}

py::array_t<uint32_t> int_decode(int codec, py::array_t<uint32_t, py::array::c_style | py::array::forcecast> w, int64_t n) {
  const uint32_t* p = w.data();
  const size_t nw = (size_t)w.size();
  switch (codec) {
    case 0: return to_np(std::vector<uint32_t>(p, p + std::min<size_t>(nw, (size_t)n)));
    case 1: return to_np(dec_vbyte(p, nw, (size_t)n));
    case 2: return to_np(dec_bp(p, nw, (size_t)n, 32));
    case 3: return to_np(dec_bp(p, nw, (size_t)n, 128));
    case 4: return to_np(dec_simple8b(p, nw, (size_t)n));
    case 5: return to_np(dec_pfor(p, nw, (size_t)n));
  }
  throw std::runtime_error("unknown integer codec id");
}

void write_csv(const std::string& path, py::array_t<double, py::array::c_style | py::array::forcecast> v) {
  FILE* f = fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("cannot open " + path);
  const double* p = v.data();
  for (py::ssize_t i = 0; i < v.size(); ++i) fprintf(f, "%.40g\n", p[i]);
  fclose(f);
}

// ---------------------------------------------------------------------------
// canonical Huffman over bytes (bitstream MSB-first, 4-byte LE symbol-count header) — the native twin of
// codecs/lossless.py::huffman_encode / huffman_decode (reference pytorch/deepreduce.py:767-802 runs the pure-Python
// `dahuffman` codec and rebuilds its model on every call)
// ---------------------------------------------------------------------------
py::array_t<uint8_t> huffman_encode(py::array_t<uint8_t, py::array::c_style | py::array::forcecast> data,
                                    py::array_t<int64_t, py::array::c_style | py::array::forcecast> lengths,
                                    py::array_t<uint64_t, py::array::c_style | py::array::forcecast> codes) {
  if (lengths.size() != 256 || codes.size() != 256) throw std::runtime_error("need 256 lengths / codes");
  const uint8_t* d = data.data();
  const int64_t* L = lengths.data();
  const uint64_t* C = codes.data();
  const size_t n = (size_t)data.size();
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    if (L[d[i]] <= 0) throw std::runtime_error("symbol without a code");
    total += (uint64_t)L[d[i]];
  }
  std::vector<uint8_t> out(4 + (total + 7) / 8, 0);
  out[0] = n & 0xFF; out[1] = (n >> 8) & 0xFF; out[2] = (n >> 16) & 0xFF; out[3] = (n >> 24) & 0xFF;
  uint64_t bit = 0;
  for (size_t i = 0; i < n; ++i) {
    const int len = (int)L[d[i]];
    const uint64_t code = C[d[i]];
    for (int b = len - 1; b >= 0; --b, ++bit)
      if ((code >> b) & 1u) out[4 + (bit >> 3)] |= (uint8_t)(0x80u >> (bit & 7));
  }
  py::array_t<uint8_t> r(out.size());
  std::memcpy(r.mutable_data(), out.data(), out.size());
  return r;
}

py::array_t<uint8_t> huffman_decode(py::array_t<uint8_t, py::array::c_style | py::array::forcecast> stream,
                                    py::array_t<int64_t, py::array::c_style | py::array::forcecast> lengths,
                                    py::array_t<uint64_t, py::array::c_style | py::array::forcecast> codes) {
  if (lengths.size() != 256 || codes.size() != 256) throw std::runtime_error("need 256 lengths / codes");
  if (stream.size() < 4) throw std::runtime_error("truncated huffman stream");
  const uint8_t* s = stream.data();
  const size_t n = (size_t)s[0] | ((size_t)s[1] << 8) | ((size_t)s[2] << 16) | ((size_t)s[3] << 24);
  // canonical decoding: per length, first code and the symbols of that length in code order
  int max_len = 0;
  for (int i = 0; i < 256; ++i) max_len = std::max<int>(max_len, (int)lengths.data()[i]);
  std::vector<std::vector<std::pair<uint64_t, uint8_t>>> by_len(max_len + 1);
  for (int i = 0; i < 256; ++i)
    if (lengths.data()[i] > 0) by_len[lengths.data()[i]].push_back({codes.data()[i], (uint8_t)i});
  for (auto& v : by_len) std::sort(v.begin(), v.end());
  py::array_t<uint8_t> out(n);
  uint8_t* o = out.mutable_data();
  const uint64_t total_bits = (uint64_t)(stream.size() - 4) * 8;
  uint64_t bit = 0, code = 0;
  int len = 0;
  size_t j = 0;
  while (j < n) {
    if (bit >= total_bits) throw std::runtime_error("huffman stream ended early");
    code = (code << 1) | ((s[4 + (bit >> 3)] >> (7 - (bit & 7))) & 1u);
    ++bit; ++len;
    if (len > max_len) throw std::runtime_error("invalid huffman code");
    const auto& v = by_len[len];
    if (!v.empty() && code >= v.front().first && code <= v.back().first) {
      const size_t k = (size_t)(code - v.front().first);          // canonical: codes of one length are consecutive
      if (k < v.size() && v[k].first == code) { o[j++] = v[k].second; code = 0; len = 0; }
    }
  }
  return out;
}

// ---------------------------------------------------------------------------
// Bloom filter object for Python (the surface the reference's ops use of bloom::OrdinaryBloomFilter,
// reference tensorflow/bloom_filter_compression.cc:102-110,206: insert / query / raw words / byte and hash counts /
// hash positions / false-positive count over a universe / construction from received raw words)
// ---------------------------------------------------------------------------
struct PyBloom {
  Bloom b;
  PyBloom(uint32_t k, uint32_t m_bits, uint32_t seed) : b(k, m_bits, seed) {}
  static PyBloom from_words(py::array_t<uint32_t, py::array::c_style | py::array::forcecast> w, uint32_t k,
                            uint32_t m_bits, uint32_t seed) {
    PyBloom f(k, m_bits, seed);
    if ((size_t)w.size() != f.b.words.size()) throw std::runtime_error("word count does not match m_bits");
    std::copy(w.data(), w.data() + w.size(), f.b.words.begin());
    return f;
  }
  void insert(py::array_t<int64_t, py::array::c_style | py::array::forcecast> idx) {
    for (py::ssize_t i = 0; i < idx.size(); ++i) b.insert((uint32_t)idx.data()[i]);
  }
  py::array_t<bool> query(py::array_t<int64_t, py::array::c_style | py::array::forcecast> idx) const {
    py::array_t<bool> out(idx.size());
    bool* o = out.mutable_data();
    for (py::ssize_t i = 0; i < idx.size(); ++i) o[i] = b.query((uint32_t)idx.data()[i]);
    return out;
  }
  py::array_t<uint32_t> words() const { return to_np(b.words); }
  size_t num_bytes() const { return b.words.size() * 4; }
  uint32_t num_hashes() const { return b.k; }
  uint32_t hash(uint32_t x, uint32_t j) const { return b.pos(x, j); }
  // positives in [0, N) that are not in the (sorted or unsorted) inserted set
  int64_t compute_false_positives(int64_t N, py::array_t<int64_t, py::array::c_style | py::array::forcecast> inserted) const {
    std::vector<int64_t> t(inserted.data(), inserted.data() + inserted.size());
    std::sort(t.begin(), t.end());
    int64_t fp = 0;
    for (int64_t x = 0; x < N; ++x)
      if (b.query((uint32_t)x) && !std::binary_search(t.begin(), t.end(), x)) ++fp;
    return fp;
  }
};

}  // namespace

PYBIND11_MODULE(_dr_cpu, m) {
  m.doc() = "DeepReduce-B200 host-side native ops";
  py::class_<PyBloom>(m, "BloomFilter")
      .def(py::init<uint32_t, uint32_t, uint32_t>(), py::arg("num_hashes"), py::arg("m_bits"), py::arg("seed"))
      .def_static("from_words", &PyBloom::from_words)
      .def("insert", &PyBloom::insert)
      .def("query", &PyBloom::query)
      .def("words", &PyBloom::words)
      .def("num_bytes", &PyBloom::num_bytes)
      .def("num_hashes", &PyBloom::num_hashes)
      .def("hash", &PyBloom::hash)
      .def("compute_false_positives", &PyBloom::compute_false_positives);
  m.def("bloom_insert", &bloom_insert);
  m.def("bloom_select", &bloom_select);
  m.def("conflict_sets", &conflict_sets);
  m.def("int_encode", &int_encode);
  m.def("int_decode", &int_decode);
  m.def("write_csv", &write_csv);
  m.def("huffman_encode", &huffman_encode);
  m.def("huffman_decode", &huffman_decode);
}
