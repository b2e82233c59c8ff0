// CPU-only native helpers of sparkflow_b200 (module sparkflow_b200._host):
//   * crc32c (Castagnoli, slice-by-8) + TF/leveldb masking
//   * leveldb-format SSTable reader / writer - the container of TensorFlow's V2 checkpoint ".index"
//     (prefix-compressed keys, restart arrays, block trailers with masked crc32c, 48-byte footer)
//   * the StopWordsRemover carrier codec (bytes <-> "b0,b1,...," decimal text)
//   * a numeric CSV parser (mnist_train.csv: 42,000 x 785) and a threaded row gather for minibatches
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace py = pybind11;

namespace {

// ------------------------------------------------------------------------------------------------
// crc32c
// ------------------------------------------------------------------------------------------------
struct Crc32cTable {
  uint32_t t[8][256];
  Crc32cTable() {
    const uint32_t poly = 0x82F63B78u;
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ poly : (c >> 1);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};
const Crc32cTable& crc_table() {
  static Crc32cTable tbl;
  return tbl;
}
uint32_t crc32c_extend(uint32_t crc, const uint8_t* p, size_t n) {
  const auto& T = crc_table().t;
  uint32_t c = crc ^ 0xFFFFFFFFu;
  while (n >= 8) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    v ^= c;
    c = T[7][v & 0xFF] ^ T[6][(v >> 8) & 0xFF] ^ T[5][(v >> 16) & 0xFF] ^ T[4][(v >> 24) & 0xFF] ^
        T[3][(v >> 32) & 0xFF] ^ T[2][(v >> 40) & 0xFF] ^ T[1][(v >> 48) & 0xFF] ^ T[0][(v >> 56) & 0xFF];
    p += 8;
    n -= 8;
  }
  while (n--) c = T[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
constexpr uint32_t kMaskDelta = 0xa282ead8u;
inline uint32_t crc_mask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + kMaskDelta; }
inline uint32_t crc_unmask(uint32_t m) {
  const uint32_t rot = m - kMaskDelta;
  return (rot >> 17) | (rot << 15);
}

// ------------------------------------------------------------------------------------------------
// varints / fixed ints
// ------------------------------------------------------------------------------------------------
void put_varint(std::string& s, uint64_t v) {
  while (v >= 128) {
    s.push_back(static_cast<char>(v | 128));
    v >>= 7;
  }
  s.push_back(static_cast<char>(v));
}
uint64_t get_varint(const uint8_t*& p, const uint8_t* end) {
  uint64_t r = 0;
  for (int shift = 0; shift <= 63 && p < end; shift += 7) {
    const uint64_t b = *p++;
    r |= (b & 127) << shift;
    if (!(b & 128)) return r;
  }
  throw std::runtime_error("sstable: malformed varint");
}
void put_fixed32(std::string& s, uint32_t v) { s.append(reinterpret_cast<const char*>(&v), 4); }
uint32_t get_fixed32(const uint8_t* p) {
  uint32_t v;
  std::memcpy(&v, p, 4);
  return v;
}

constexpr uint64_t kTableMagic = 0xdb4775248b80fb57ull;
constexpr size_t kFooterSize = 48;
constexpr size_t kTrailerSize = 5;

struct Handle {
  uint64_t offset = 0, size = 0;
};

// ------------------------------------------------------------------------------------------------
// SSTable reader
// ------------------------------------------------------------------------------------------------
std::string read_block(const std::string& file, const Handle& h, bool verify) {
  if (h.offset + h.size + kTrailerSize > file.size()) throw std::runtime_error("sstable: block handle out of range");
  const uint8_t* base = reinterpret_cast<const uint8_t*>(file.data()) + h.offset;
  const uint8_t type = base[h.size];
  if (verify) {
    const uint32_t stored = crc_unmask(get_fixed32(base + h.size + 1));
    const uint32_t actual = crc32c_extend(0, base, h.size + 1);
    if (stored != actual) throw std::runtime_error("sstable: block checksum mismatch");
  }
  if (type != 0) throw std::runtime_error("sstable: compressed blocks are not supported (TF bundles are uncompressed)");
  return std::string(reinterpret_cast<const char*>(base), h.size);
}

void parse_block(const std::string& blk, std::vector<std::pair<std::string, std::string>>& out) {
  if (blk.size() < 4) throw std::runtime_error("sstable: short block");
  const uint32_t num_restarts = get_fixed32(reinterpret_cast<const uint8_t*>(blk.data()) + blk.size() - 4);
  const size_t data_end = blk.size() - 4 - static_cast<size_t>(num_restarts) * 4;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(blk.data());
  const uint8_t* end = p + data_end;
  std::string key;
  while (p < end) {
    const uint64_t shared = get_varint(p, end), non_shared = get_varint(p, end), vlen = get_varint(p, end);
    if (shared > key.size() || p + non_shared + vlen > end) throw std::runtime_error("sstable: corrupt entry");
    key.resize(shared);
    key.append(reinterpret_cast<const char*>(p), non_shared);
    p += non_shared;
    out.emplace_back(key, std::string(reinterpret_cast<const char*>(p), vlen));
    p += vlen;
  }
}

py::list sstable_read(const std::string& file, bool verify) {
  if (file.size() < kFooterSize) throw std::runtime_error("sstable: file too short");
  const uint8_t* f = reinterpret_cast<const uint8_t*>(file.data()) + file.size() - kFooterSize;
  uint64_t magic;
  std::memcpy(&magic, f + 40, 8);
  if (magic != kTableMagic) throw std::runtime_error("sstable: bad magic number (not a TF checkpoint index)");
  const uint8_t* p = f;
  Handle meta, index;
  meta.offset = get_varint(p, f + 40);
  meta.size = get_varint(p, f + 40);
  index.offset = get_varint(p, f + 40);
  index.size = get_varint(p, f + 40);
  std::vector<std::pair<std::string, std::string>> idx_entries, entries;
  parse_block(read_block(file, index, verify), idx_entries);
  for (auto& ie : idx_entries) {
    const uint8_t* hp = reinterpret_cast<const uint8_t*>(ie.second.data());
    const uint8_t* he = hp + ie.second.size();
    Handle h;
    h.offset = get_varint(hp, he);
    h.size = get_varint(hp, he);
    parse_block(read_block(file, h, verify), entries);
  }
  py::list out;
  for (auto& e : entries) out.append(py::make_tuple(py::bytes(e.first), py::bytes(e.second)));
  return out;
}

// ------------------------------------------------------------------------------------------------
// SSTable writer (keys must arrive sorted)
// ------------------------------------------------------------------------------------------------
struct BlockBuilder {
  std::string buf, last_key;
  std::vector<uint32_t> restarts{0};
  int counter = 0;
  static constexpr int kRestartInterval = 16;
  void add(const std::string& key, const std::string& value) {
    size_t shared = 0;
    if (counter < kRestartInterval) {
      const size_t m = std::min(last_key.size(), key.size());
      while (shared < m && last_key[shared] == key[shared]) ++shared;
    } else {
      restarts.push_back(static_cast<uint32_t>(buf.size()));
      counter = 0;
    }
    put_varint(buf, shared);
    put_varint(buf, key.size() - shared);
    put_varint(buf, value.size());
    buf.append(key, shared, std::string::npos);
    buf.append(value);
    last_key = key;
    ++counter;
  }
  std::string finish() {
    std::string out = buf;
    for (uint32_t r : restarts) put_fixed32(out, r);
    put_fixed32(out, static_cast<uint32_t>(restarts.size()));
    return out;
  }
  bool empty() const { return buf.empty(); }
};

Handle append_block(std::string& file, const std::string& contents) {
  Handle h;
  h.offset = file.size();
  h.size = contents.size();
  file.append(contents);
  const char type = 0;
  file.push_back(type);
  uint32_t crc = crc32c_extend(0, reinterpret_cast<const uint8_t*>(contents.data()), contents.size());
  crc = crc32c_extend(crc, reinterpret_cast<const uint8_t*>(&type), 1);
  put_fixed32(file, crc_mask(crc));
  return h;
}

// shortest key k with  start <= k < limit  (leveldb BytewiseComparator::FindShortestSeparator)
std::string shortest_separator(const std::string& start, const std::string& limit) {
  const size_t m = std::min(start.size(), limit.size());
  size_t d = 0;
  while (d < m && start[d] == limit[d]) ++d;
  if (d >= m) return start;
  const uint8_t b = static_cast<uint8_t>(start[d]);
  if (b < 0xff && b + 1 < static_cast<uint8_t>(limit[d])) {
    std::string r = start.substr(0, d + 1);
    r[d] = static_cast<char>(b + 1);
    return r;
  }
  return start;
}
std::string short_successor(const std::string& key) {
  for (size_t i = 0; i < key.size(); ++i) {
    const uint8_t b = static_cast<uint8_t>(key[i]);
    if (b != 0xff) {
      std::string r = key.substr(0, i + 1);
      r[i] = static_cast<char>(b + 1);
      return r;
    }
  }
  return key;
}

py::bytes sstable_write(const std::vector<std::pair<std::string, std::string>>& entries, size_t block_size) {
  for (size_t i = 1; i < entries.size(); ++i)
    if (!(entries[i - 1].first < entries[i].first)) throw std::runtime_error("sstable: keys must be strictly increasing");
  std::string file;
  BlockBuilder data, index;
  std::string pending_last_key;
  bool pending = false;
  Handle pending_handle;
  auto flush = [&](const std::string* next_key) {
    if (data.empty()) return;
    pending_last_key = data.last_key;
    pending_handle = append_block(file, data.finish());
    data = BlockBuilder();
    pending = true;
    if (pending) {
      std::string sep = next_key ? shortest_separator(pending_last_key, *next_key) : short_successor(pending_last_key);
      std::string hv;
      put_varint(hv, pending_handle.offset);
      put_varint(hv, pending_handle.size);
      index.add(sep, hv);
      pending = false;
    }
  };
  for (size_t i = 0; i < entries.size(); ++i) {
    data.add(entries[i].first, entries[i].second);
    if (data.buf.size() >= block_size) flush(i + 1 < entries.size() ? &entries[i + 1].first : nullptr);
  }
  flush(nullptr);
  BlockBuilder metaindex;
  const Handle mh = append_block(file, metaindex.finish());
  const Handle ih = append_block(file, index.finish());
  std::string footer;
  put_varint(footer, mh.offset);
  put_varint(footer, mh.size);
  put_varint(footer, ih.offset);
  put_varint(footer, ih.size);
  footer.resize(40, '\0');
  footer.append(reinterpret_cast<const char*>(&kTableMagic), 8);
  file.append(footer);
  return py::bytes(file);
}

// ------------------------------------------------------------------------------------------------
// carrier codec
// ------------------------------------------------------------------------------------------------
std::string bytes_to_decimal_csv(const std::string& raw) {
  static const char digits[] = "0123456789";
  std::string out;
  out.reserve(raw.size() * 4);
  for (unsigned char b : raw) {
    if (b >= 100) {
      out.push_back(digits[b / 100]);
      out.push_back(digits[(b / 10) % 10]);
    } else if (b >= 10) {
      out.push_back(digits[b / 10]);
    }
    out.push_back(digits[b % 10]);
    out.push_back(',');
  }
  return out;
}
py::bytes decimal_csv_to_bytes(const std::string& text) {
  std::string out;
  out.reserve(text.size() / 2);
  int cur = -1;
  for (char c : text) {
    if (c >= '0' && c <= '9') {
      cur = (cur < 0 ? 0 : cur) * 10 + (c - '0');
      if (cur > 255) throw std::runtime_error("carrier payload: byte value out of range");
    } else if (c == ',') {
      if (cur < 0) throw std::runtime_error("carrier payload: empty field");
      out.push_back(static_cast<char>(cur));
      cur = -1;
    } else if (c != ' ' && c != '\n') {
      throw std::runtime_error("carrier payload: unexpected character");
    }
  }
  // like the reference decoder, whatever follows the last comma is dropped
  return py::bytes(out);
}

// ------------------------------------------------------------------------------------------------
// CSV
// ------------------------------------------------------------------------------------------------
py::array_t<double> read_csv(const std::string& path, int skip_rows) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::string buf(static_cast<size_t>(sz), '\0');
  if (sz > 0 && std::fread(&buf[0], 1, static_cast<size_t>(sz), f) != static_cast<size_t>(sz)) {
    std::fclose(f);
    throw std::runtime_error("short read on " + path);
  }
  std::fclose(f);
  std::vector<double> vals;
  vals.reserve(static_cast<size_t>(sz) / 2);
  size_t rows = 0, cols = 0, cur_cols = 0;
  const char* p = buf.data();
  const char* end = p + buf.size();
  int to_skip = skip_rows;
  while (p < end) {
    const char* line_end = static_cast<const char*>(std::memchr(p, '\n', static_cast<size_t>(end - p)));
    if (!line_end) line_end = end;
    if (to_skip > 0) {
      --to_skip;
      p = line_end + 1;
      continue;
    }
    const char* q = p;
    cur_cols = 0;
    while (q < line_end) {
      // fast path: plain (signed) integers, which is all of mnist_train.csv
      {
        const char* r = q;
        bool neg = false;
        if (r < line_end && (*r == '-' || *r == '+')) { neg = (*r == '-'); ++r; }
        const char* d0 = r;
        uint64_t acc = 0;
        while (r < line_end && *r >= '0' && *r <= '9' && r - d0 < 18) { acc = acc * 10 + static_cast<uint64_t>(*r - '0'); ++r; }
        if (r > d0 && (r == line_end || *r == ',' || *r == '\r')) {
          vals.push_back(neg ? -static_cast<double>(acc) : static_cast<double>(acc));
          ++cur_cols;
          q = (r < line_end && *r == ',') ? r + 1 : line_end;
          continue;
        }
      }
      char* next = nullptr;
      const double v = std::strtod(q, &next);
      if (next == q) {               // empty / non-numeric field -> NaN
        vals.push_back(std::nan(""));
        while (q < line_end && *q != ',') ++q;
      } else {
        vals.push_back(v);
        q = next;
      }
      ++cur_cols;
      while (q < line_end && *q != ',') ++q;
      if (q < line_end) ++q;
    }
    if (cur_cols > 0) {
      if (cols == 0) cols = cur_cols;
      if (cur_cols != cols) throw std::runtime_error("csv: ragged row " + std::to_string(rows));
      ++rows;
    }
    p = line_end + 1;
  }
  py::array_t<double> arr({rows, cols});
  if (rows * cols) std::memcpy(arr.mutable_data(), vals.data(), rows * cols * sizeof(double));
  return arr;
}

// ------------------------------------------------------------------------------------------------
// minibatch row gather (host side of the feeder): dst[i, :] = src[idx[i], :]
// ------------------------------------------------------------------------------------------------
void gather_rows(uintptr_t src, uintptr_t dst, py::array_t<int64_t, py::array::c_style | py::array::forcecast> idx,
                 size_t row_bytes, int threads) {
  const int64_t* ip = idx.data();
  const size_t n = static_cast<size_t>(idx.size());
  const char* s = reinterpret_cast<const char*>(src);
  char* d = reinterpret_cast<char*>(dst);
  py::gil_scoped_release nogil;
  auto work = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) std::memcpy(d + i * row_bytes, s + static_cast<size_t>(ip[i]) * row_bytes, row_bytes);
  };
  if (threads <= 1 || n < 64) {
    work(0, n);
    return;
  }
  std::vector<std::thread> pool;
  const size_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
}

}  // namespace

PYBIND11_MODULE(_host, m) {
  m.doc() = "sparkflow_b200 CPU-only native helpers";
  m.def("crc32c", [](const std::string& data, uint32_t init) {
    return crc32c_extend(init, reinterpret_cast<const uint8_t*>(data.data()), data.size());
  }, py::arg("data"), py::arg("init") = 0);
  m.def("crc32c_mask", &crc_mask);
  m.def("crc32c_unmask", &crc_unmask);
  m.def("sstable_read", &sstable_read, py::arg("file"), py::arg("verify_checksums") = true);
  m.def("sstable_write", &sstable_write, py::arg("entries"), py::arg("block_size") = 4096);
  m.def("bytes_to_decimal_csv", [](const std::string& raw) { return bytes_to_decimal_csv(raw); });
  m.def("decimal_csv_to_bytes", &decimal_csv_to_bytes);
  m.def("read_csv", &read_csv, py::arg("path"), py::arg("skip_rows") = 0);
  m.def("gather_rows", &gather_rows, py::arg("src"), py::arg("dst"), py::arg("idx"), py::arg("row_bytes"), py::arg("threads") = 1);
}
