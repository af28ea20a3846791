#include "h5lite.h"

#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <ctime>
#include <stdexcept>

namespace {
constexpr uint64_t UNDEF = ~0ull;
constexpr unsigned LEAF_K = 4, INTERNAL_K = 16, CHUNK_K = 32;

struct Buf {
  std::vector<uint8_t> b;
  void u8(unsigned v) { b.push_back(uint8_t(v)); }
  void u16(unsigned v) {
    for (int i = 0; i < 2; ++i) b.push_back(uint8_t(v >> (8 * i)));
  }
  void u32(uint32_t v) {
    for (int i = 0; i < 4; ++i) b.push_back(uint8_t(v >> (8 * i)));
  }
  void u64(uint64_t v) {
    for (int i = 0; i < 8; ++i) b.push_back(uint8_t(v >> (8 * i)));
  }
  void raw(const void *p, size_t n) { b.insert(b.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  void pad8() { zeros((8 - b.size() % 8) % 8); }
};

// header message: type, size of the body (a multiple of 8), flags, 3 reserved bytes, body
void message(Buf &h, unsigned type, unsigned flags, Buf body) {
  body.pad8();
  h.u16(type);
  h.u16(unsigned(body.b.size()));
  h.u8(flags);
  h.zeros(3);
  h.raw(body.b.data(), body.b.size());
}

// version 1 object header around `nmsg` messages
std::vector<uint8_t> objectHeader(unsigned nmsg, const Buf &messages) {
  Buf h;
  h.u8(1);
  h.u8(0);
  h.u16(nmsg);
  h.u32(1);                              // object reference count
  h.u32(uint32_t(messages.b.size()));    // header size
  h.zeros(4);                            // alignment of the first message to 8 bytes
  h.raw(messages.b.data(), messages.b.size());
  return h.b;
}

size_t align8(size_t v) { return (v + 7) / 8 * 8; }
}  // namespace

H5LiteFile::H5LiteFile(const std::string &path) : _path(path) {
  _f = std::fopen(path.c_str(), "wb+");
  if (!_f) throw std::runtime_error("Error opening HDF5 file '" + path + "'.");
  flush();
}

H5LiteFile::~H5LiteFile() {
  if (_f) {
    try {
      flush();
    } catch (const std::exception &) {
    }
    std::fclose(_f);
  }
}

void H5LiteFile::put(uint64_t addr, const std::vector<uint8_t> &bytes) {
  if (std::fseek(_f, long(addr), SEEK_SET) != 0 || std::fwrite(bytes.data(), 1, bytes.size(), _f) != bytes.size())
    throw std::runtime_error("Error writing HDF5 file '" + _path + "'.");
}

void H5LiteFile::addDataset(const std::string &name, const std::vector<uint64_t> &dims, int elem_size, const void *data) {
  if (elem_size != 4 && elem_size != 8) throw std::runtime_error("Unsupported output type");
  if (name.empty() || name.find('/') != std::string::npos) throw std::runtime_error("H5LiteFile: bad dataset name '" + name + "'");
  for (const auto &e : _entries)
    if (e.name == name) throw std::runtime_error("Dataset '" + name + "' already exists in HDF5 file.");
  const unsigned rank = unsigned(dims.size());
  uint64_t count = 1;
  for (auto d : dims) count *= d;
  const uint64_t nbytes = count * uint64_t(elem_size);

  // ---- the single chunk, deflate level 9 (H5Pset_deflate(plist, 9))
  uLongf zlen = compressBound(uLong(nbytes));
  std::vector<uint8_t> z(zlen);
  if (compress2(z.data(), &zlen, (const Bytef *)data, uLong(nbytes), 9) != Z_OK) throw std::runtime_error("H5LiteFile: deflate failed");
  z.resize(zlen);
  const uint64_t chunk_addr = _data_end;
  put(chunk_addr, z);

  // ---- chunk B-tree: one leaf node with one entry, allocated at its full size (2 K + 1 keys, 2 K children)
  const uint64_t tree_addr = align8(chunk_addr + z.size());
  const size_t key_size = 8 + 8 * (rank + 1);
  Buf t;
  t.raw("TREE", 4);
  t.u8(1);   // node type: raw data chunks
  t.u8(0);   // level
  t.u16(1);  // entries used
  t.u64(UNDEF);
  t.u64(UNDEF);
  t.u32(uint32_t(z.size()));  // key 0: chunk size after the filters, filter mask, chunk offset (+ 0 for the element dimension)
  t.u32(0);
  for (unsigned i = 0; i <= rank; ++i) t.u64(0);
  t.u64(chunk_addr);
  t.u32(0);  // key 1: the offset just past the dataset (libhdf5 puts the element size in the extra dimension)
  t.u32(0);
  for (unsigned i = 0; i < rank; ++i) t.u64(dims[i]);
  t.u64(uint64_t(elem_size));
  t.zeros(24 + 2 * CHUNK_K * 8 + (2 * CHUNK_K + 1) * key_size - t.b.size());
  put(tree_addr, t.b);

  // ---- object header: dataspace, datatype, fill value, filter pipeline, layout, modification time
  Buf m;
  {
    Buf s;  // dataspace version 1 with maximum dimensions = dimensions
    s.u8(1);
    s.u8(rank);
    s.u8(1);
    s.zeros(5);
    for (auto d : dims) s.u64(d);
    for (auto d : dims) s.u64(d);
    message(m, 0x0001, 0, s);
  }
  {
    Buf d;  // IEEE little-endian floating point, version 1
    d.u8(0x11);
    d.u8(0x20);                       // little endian, mantissa normalisation: implied leading 1
    d.u8(elem_size == 8 ? 63 : 31);   // sign bit
    d.u8(0);
    d.u32(uint32_t(elem_size));
    d.u16(0);                          // bit offset
    d.u16(elem_size * 8);              // precision
    d.u8(elem_size == 8 ? 52 : 23);    // exponent location
    d.u8(elem_size == 8 ? 11 : 8);     // exponent size
    d.u8(0);                           // mantissa location
    d.u8(elem_size == 8 ? 52 : 23);    // mantissa size
    d.u32(elem_size == 8 ? 1023 : 127);
    message(m, 0x0003, 1, d);
  }
  {
    Buf f;  // fill value version 2: incremental allocation, written if set, undefined
    f.u8(2);
    f.u8(3);
    f.u8(2);
    f.u8(1);
    f.u32(0);
    message(m, 0x0005, 1, f);
  }
  {
    Buf p;  // filter pipeline version 1: deflate (id 1), optional, one client value (the level)
    p.u8(1);
    p.u8(1);
    p.zeros(6);
    p.u16(1);
    p.u16(8);
    p.u16(1);
    p.u16(1);
    p.raw("deflate\0", 8);
    p.u32(9);
    p.u32(0);
    message(m, 0x000B, 1, p);
  }
  {
    Buf l;  // layout version 3, chunked: dimensionality rank + 1, B-tree address, chunk dimensions, element size
    l.u8(3);
    l.u8(2);
    l.u8(rank + 1);
    l.u64(tree_addr);
    for (auto d : dims) l.u32(uint32_t(d));
    l.u32(uint32_t(elem_size));
    message(m, 0x0008, 0, l);
  }
  {
    Buf mt;  // object modification time, version 1
    mt.u8(1);
    mt.zeros(3);
    mt.u32(uint32_t(std::time(nullptr)));
    message(m, 0x0012, 0, mt);
  }
  const uint64_t header_addr = align8(tree_addr + t.b.size());
  const auto hdr = objectHeader(6, m);
  put(header_addr, hdr);
  _data_end = align8(header_addr + hdr.size());
  Entry e;
  e.name = name;
  e.header = header_addr;
  _entries.push_back(e);
}

void H5LiteFile::flush() {
  // ---- local heap: "" for the root at offset 0, then the names, 8-byte aligned
  std::vector<Entry *> order;
  for (auto &e : _entries) order.push_back(&e);
  std::sort(order.begin(), order.end(), [](const Entry *a, const Entry *b) { return std::strcmp(a->name.c_str(), b->name.c_str()) < 0; });
  Buf heap_data;
  heap_data.zeros(8);
  for (auto &e : _entries) {  // creation order, like libhdf5
    e.heap_off = heap_data.b.size();
    heap_data.raw(e.name.c_str(), e.name.size() + 1);
    heap_data.pad8();
  }
  // a free block at the end keeps the heap's free list well-formed (next = 1: none; size of the block)
  const uint64_t free_off = heap_data.b.size();
  heap_data.u64(1);
  heap_data.u64(16);

  uint64_t cur = _data_end;
  auto alloc = [&](size_t n) {
    const uint64_t a = cur;
    cur = align8(cur + n);
    return a;
  };
  const uint64_t root_hdr = alloc(16 + 8 + 16);
  const uint64_t heap_hdr = alloc(32);
  const uint64_t heap_dat = alloc(heap_data.b.size());

  // ---- symbol table nodes (up to 2 LEAF_K entries, sorted by name) under B-tree nodes of up to 2 INTERNAL_K children
  struct Node {
    uint64_t addr, last_key;  // address and the heap offset of the largest name below it
  };
  const size_t snod_size = 8 + 2 * LEAF_K * 40, tree_size = 24 + 2 * INTERNAL_K * 8 + (2 * INTERNAL_K + 1) * 8;
  std::vector<Node> level;
  std::vector<std::pair<uint64_t, std::vector<uint8_t>>> blocks;
  const size_t per_leaf = 2 * LEAF_K;
  for (size_t i = 0; i < order.size() || (i == 0 && order.empty()); i += per_leaf) {
    const size_t n = std::min(per_leaf, order.size() - i);
    Buf s;
    s.raw("SNOD", 4);
    s.u8(1);
    s.u8(0);
    s.u16(unsigned(n));
    for (size_t k = 0; k < n; ++k) {
      s.u64(order[i + k]->heap_off);
      s.u64(order[i + k]->header);
      s.u32(0);  // cache type: nothing cached
      s.u32(0);
      s.zeros(16);
    }
    s.zeros(snod_size - s.b.size());
    const uint64_t a = alloc(snod_size);
    blocks.push_back({a, s.b});
    level.push_back({a, n ? order[i + n - 1]->heap_off : 0});
    if (order.empty()) break;
  }
  unsigned depth = 0;
  while (true) {
    std::vector<Node> up;
    for (size_t i = 0; i < level.size(); i += 2 * INTERNAL_K) {
      const size_t n = std::min<size_t>(2 * INTERNAL_K, level.size() - i);
      Buf t;
      t.raw("TREE", 4);
      t.u8(0);  // node type: group
      t.u8(depth);
      t.u16(unsigned(n));
      t.u64(UNDEF);  // siblings: filled below for the nodes of one level
      t.u64(UNDEF);
      t.u64(i == 0 ? 0 : level[i - 1].last_key);
      for (size_t k = 0; k < n; ++k) {
        t.u64(level[i + k].addr);
        t.u64(level[i + k].last_key);
      }
      t.zeros(tree_size - t.b.size());
      const uint64_t a = alloc(tree_size);
      blocks.push_back({a, t.b});
      up.push_back({a, level[i + n - 1].last_key});
    }
    // sibling links within the level
    for (size_t j = 0; j < up.size(); ++j) {
      auto &blk = blocks[blocks.size() - up.size() + j].second;
      const uint64_t left = j ? up[j - 1].addr : UNDEF, right = j + 1 < up.size() ? up[j + 1].addr : UNDEF;
      std::memcpy(&blk[8], &left, 8);
      std::memcpy(&blk[16], &right, 8);
    }
    level = up;
    ++depth;
    if (level.size() == 1) break;
  }
  const uint64_t btree = level[0].addr;

  for (const auto &b : blocks) put(b.first, b.second);
  put(heap_dat, heap_data.b);
  {
    Buf h;
    h.raw("HEAP", 4);
    h.u8(0);
    h.zeros(3);
    h.u64(heap_data.b.size());
    h.u64(free_off);
    h.u64(heap_dat);
    put(heap_hdr, h.b);
  }
  {
    Buf m, st;
    st.u64(btree);
    st.u64(heap_hdr);
    message(m, 0x0011, 0, st);
    put(root_hdr, objectHeader(1, m));
  }
  const uint64_t eof = cur;
  {
    Buf s;  // superblock version 0
    s.raw("\x89HDF\r\n\x1a\n", 8);
    s.u8(0);  // superblock, free-space storage, root group symbol table entry versions
    s.u8(0);
    s.u8(0);
    s.u8(0);
    s.u8(0);  // shared header message format version
    s.u8(8);  // size of offsets, size of lengths
    s.u8(8);
    s.u8(0);
    s.u16(LEAF_K);
    s.u16(INTERNAL_K);
    s.u32(0);  // file consistency flags
    s.u64(0);  // base address
    s.u64(UNDEF);
    s.u64(eof);
    s.u64(UNDEF);
    s.u64(0);  // root group symbol table entry: name offset, object header, cache type 1 with the B-tree / heap addresses
    s.u64(root_hdr);
    s.u32(1);
    s.u32(0);
    s.u64(btree);
    s.u64(heap_hdr);
    put(0, s.b);
  }
  std::fflush(_f);
  // drop what an earlier, longer metadata block may have left behind the new end of file
  if (std::fseek(_f, 0, SEEK_END) == 0 && uint64_t(std::ftell(_f)) > eof) {
    if (ftruncate(fileno(_f), off_t(eof)) != 0) throw std::runtime_error("Error truncating HDF5 file '" + _path + "'.");
  } else if (uint64_t(std::ftell(_f)) < eof) {
    put(eof - 1, std::vector<uint8_t>(1, 0));
    std::fflush(_f);
  }
}
