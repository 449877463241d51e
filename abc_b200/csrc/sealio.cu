// sealio.cu — Microsoft SEAL 3.6 binary streams <-> device handles (host code only; no kernels here).
//
// north_star: "bit-exact with SEAL for the same SEAL-serialised input ciphertexts and keys".  ABC itself never
// serialises (its tests encrypt in-process), so this is the door through which artefacts of a real
// SEAL 3.6.5 process enter the CUDA backend and leave it again: seal::Ciphertext, SecretKey, PublicKey, RelinKeys,
// GaloisKeys and EncryptionParameters, as written by seal::Serialization::Save
// (SEAL 3.6.5 native/src/seal/serialization.{h,cpp}, ciphertext.cpp, plaintext.cpp, publickey.h, secretkey.h,
// kswitchkeys.cpp, encryptionparams.cpp, modulus.cpp, dynarray.h — restated from the published format, SEAL is not
// in /root/reference).  "parity unpinned": no SEAL-written byte stream exists in this environment; the format is
// checked against an independent Python restatement (tests/seal_format.py) and BLAKE2b against hashlib.
//
// Record = 16-byte header {u16 magic 0xA15E, u8 header_size 0x10, u8 version_major, u8 version_minor,
// u8 compr_mode (0 none, 1 zlib, 2 zstd), u16 reserved, u64 size (header included)} + body; with a compr_mode the body
// is ONE deflate / zstd stream.  Nested records (DynArray, Plaintext, Ciphertext inside keys, Modulus) are written with
// compr_mode none and carry their own header.
#include <dlfcn.h>
#include <zlib.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/abc_b200.h"
#include "hostmath.hpp"

extern "C" void abc_set_error(abc_ctx *ctx, const char *msg);  // lib.cu

namespace {
typedef unsigned long long u64;

// ---------------------------------------------------------------- BLAKE2b (RFC 7693), unkeyed, for parms_id
// util/hash.h HashFunction::hash = blake2b(out, 32, in, 8 * count, key = none)
const u64 B2_IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                      0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
const unsigned char B2_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
inline u64 rotr(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
void b2_compress(u64 h[8], const unsigned char block[128], u64 t, bool last) {
  u64 m[16], v[16];
  memcpy(m, block, 128);  // little-endian host
  for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = B2_IV[i]; }
  v[12] ^= t;
  if (last) v[14] = ~v[14];
  for (int r = 0; r < 12; ++r) {
    const unsigned char *s = B2_SIGMA[r];
#define B2G(a, b, c, d, x, y)                                   \
  v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32);       \
  v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 24);       \
  v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16);       \
  v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 63);
    B2G(0, 4, 8, 12, m[s[0]], m[s[1]]) B2G(1, 5, 9, 13, m[s[2]], m[s[3]])
    B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
    B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) B2G(1, 6, 11, 12, m[s[10]], m[s[11]])
    B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
#undef B2G
  }
  for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}
void blake2b_256(const void *in, size_t len, u64 out[4]) {
  u64 h[8];
  memcpy(h, B2_IV, sizeof h);
  h[0] ^= 0x01010000ULL ^ 32;  // digest length 32, no key, fanout = depth = 1
  const unsigned char *p = (const unsigned char *)in;
  u64 t = 0;
  while (len > 128) { t += 128; b2_compress(h, p, t, false); p += 128; len -= 128; }
  unsigned char last[128] = {0};
  memcpy(last, p, len);
  t += len;
  b2_compress(h, last, t, true);
  memcpy(out, h, 32);
}

// ---------------------------------------------------------------- byte streams
struct Writer {
  std::vector<unsigned char> b;
  void raw(const void *p, size_t n) { const unsigned char *c = (const unsigned char *)p; b.insert(b.end(), c, c + n); }
  void u8(unsigned v) { b.push_back((unsigned char)v); }
  void u64v(u64 v) { raw(&v, 8); }
  void f64(double v) { raw(&v, 8); }
  // nested record with compr_mode none: header placeholder, body by `fill`, then the size
  template <typename F> void record(F fill) {
    const size_t at = b.size();
    const unsigned char hdr[16] = {0x5E, 0xA1, 0x10, 3, 6, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    raw(hdr, 16);
    fill(*this);
    const u64 size = b.size() - at;
    memcpy(&b[at + 8], &size, 8);
  }
};
struct Reader {
  const unsigned char *p; size_t n, at = 0; std::string err;
  Reader(const unsigned char *p_, size_t n_) : p(p_), n(n_) {}
  bool need(size_t k) { if (!err.empty()) return false; if (n - at < k) { err = "SEAL stream truncated"; return false; } return true; }
  bool raw(void *o, size_t k) { if (!need(k)) return false; memcpy(o, p + at, k); at += k; return true; }
  u64 u64v() { u64 v = 0; raw(&v, 8); return v; }
  unsigned u8() { unsigned char v = 0; raw(&v, 1); return v; }
  double f64() { double v = 0; raw(&v, 8); return v; }
  // header of a nested (uncompressed) record; returns its declared size
  u64 header_plain() {
    unsigned char h[16];
    if (!raw(h, 16)) return 0;
    if (h[0] != 0x5E || h[1] != 0xA1 || h[2] != 0x10) { err = "not a SEAL record (magic / header size)"; return 0; }
    if (h[3] != 3) { err = "SEAL record written by an incompatible major version"; return 0; }
    if (h[5] != 0) { err = "nested SEAL record is compressed (unexpected)"; return 0; }
    u64 size; memcpy(&size, h + 8, 8);
    if (size < 16 || size - 16 > n - at) { err = "SEAL record size field out of range"; return 0; }
    return size;
  }
};

// ---------------------------------------------------------------- compression of the outermost record
typedef size_t (*zstd_compress_t)(void *, size_t, const void *, size_t, int);
typedef size_t (*zstd_bound_t)(size_t);
typedef unsigned (*zstd_iserr_t)(size_t);
struct ZBuf { const void *src; size_t size, pos; };
struct ZOut { void *dst; size_t size, pos; };
typedef void *(*zstd_cds_t)();
typedef size_t (*zstd_fds_t)(void *);
typedef size_t (*zstd_ds_t)(void *, ZOut *, ZBuf *);
struct Zstd {
  void *h = nullptr; zstd_compress_t compress; zstd_bound_t bound; zstd_iserr_t iserr; zstd_cds_t cds; zstd_fds_t fds; zstd_ds_t ds;
  bool load() {
    if (h) return true;
    h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) return false;
    compress = (zstd_compress_t)dlsym(h, "ZSTD_compress"); bound = (zstd_bound_t)dlsym(h, "ZSTD_compressBound");
    iserr = (zstd_iserr_t)dlsym(h, "ZSTD_isError"); cds = (zstd_cds_t)dlsym(h, "ZSTD_createDStream");
    fds = (zstd_fds_t)dlsym(h, "ZSTD_freeDStream"); ds = (zstd_ds_t)dlsym(h, "ZSTD_decompressStream");
    return compress && bound && iserr && cds && fds && ds;
  }
} g_zstd;

bool deflate_all(const std::vector<unsigned char> &in, std::vector<unsigned char> &out) {
  uLongf cap = compressBound((uLong)in.size());
  out.resize(cap);
  if (compress2(out.data(), &cap, in.data(), (uLong)in.size(), Z_DEFAULT_COMPRESSION) != Z_OK) return false;
  out.resize(cap);
  return true;
}
// Streams are untrusted input: the decompressed size is capped (max_out: the largest record this context can legally hold,
// see max_record_bytes) so a decompression bomb cannot exhaust host memory, and zlib is fed in pieces below UINT_MAX so
// streams of 4 GiB or more are not silently truncated.
bool inflate_all(const unsigned char *in, size_t n, std::vector<unsigned char> &out, std::string &err, size_t max_out) {
  z_stream z;
  memset(&z, 0, sizeof z);
  if (inflateInit(&z) != Z_OK) { err = "zlib init failed"; return false; }
  out.clear();
  unsigned char chunk[1 << 16];
  size_t fed = 0;
  int r = Z_OK;
  while (r != Z_STREAM_END) {
    if (z.avail_in == 0 && fed < n) {
      const size_t piece = std::min<size_t>(n - fed, (size_t)1 << 30);
      z.next_in = const_cast<unsigned char *>(in + fed); z.avail_in = (uInt)piece;
      fed += piece;
    }
    z.next_out = chunk; z.avail_out = sizeof chunk;
    r = inflate(&z, Z_NO_FLUSH);
    if (r != Z_OK && r != Z_STREAM_END) { inflateEnd(&z); err = "zlib stream corrupt"; return false; }
    out.insert(out.end(), chunk, chunk + (sizeof chunk - z.avail_out));
    if (out.size() > max_out) { inflateEnd(&z); err = "compressed SEAL record expands beyond the largest record of this parameter set"; return false; }
    if (r == Z_OK && z.avail_in == 0 && fed == n && z.avail_out != 0) { inflateEnd(&z); err = "zlib stream truncated"; return false; }
  }
  inflateEnd(&z);
  return true;
}
bool unzstd_all(const unsigned char *in, size_t n, std::vector<unsigned char> &out, std::string &err, size_t max_out) {
  if (!g_zstd.load()) { err = "zstd-compressed SEAL stream, and libzstd.so.1 is not loadable here"; return false; }
  void *ds = g_zstd.cds();
  ZBuf ib{in, n, 0};
  out.clear();
  std::vector<unsigned char> chunk(1 << 17);
  while (ib.pos < ib.size) {
    ZOut ob{chunk.data(), chunk.size(), 0};
    const size_t r = g_zstd.ds(ds, &ob, &ib);
    if (g_zstd.iserr(r)) { g_zstd.fds(ds); err = "zstd stream corrupt"; return false; }
    out.insert(out.end(), chunk.data(), chunk.data() + ob.pos);
    if (out.size() > max_out) { g_zstd.fds(ds); err = "compressed SEAL record expands beyond the largest record of this parameter set"; return false; }
    if (r == 0 && ib.pos < ib.size) continue;  // next frame
  }
  g_zstd.fds(ds);
  return true;
}
// the largest record a context can legally load: a Galois key set (2 * log2(N) keys of L * 2 * k * N words) plus headers;
// without a context (parameter streams) a parameter record is a few hundred bytes
size_t max_record_bytes(const abc_ctx *c) {
  if (!c) return (size_t)1 << 20;
  const size_t N = abc_poly_degree(c), k = abc_n_primes(c), L = abc_n_limbs(c);
  size_t logn = 0;
  while (((size_t)1 << logn) < N) ++logn;
  return (2 * logn + 2) * (L * 2 * k * N * 8 + 4096) + ((size_t)1 << 20);
}

// outermost record: body bytes -> stream (with the requested compr_mode)
abc_status emit(abc_ctx *c, const Writer &body, int compr, uint8_t *buf, size_t cap, size_t *len) {
  std::vector<unsigned char> packed;
  const std::vector<unsigned char> *payload = &body.b;
  if (compr == ABC_SEAL_COMPR_ZLIB) {
    if (!deflate_all(body.b, packed)) { abc_set_error(c, "zlib deflate failed"); return ABC_ERR_STATE; }
    payload = &packed;
  } else if (compr == ABC_SEAL_COMPR_ZSTD) {
    if (!g_zstd.load()) { abc_set_error(c, "libzstd.so.1 is not loadable here"); return ABC_ERR_UNSUPPORTED; }
    packed.resize(g_zstd.bound(body.b.size()));
    const size_t r = g_zstd.compress(packed.data(), packed.size(), body.b.data(), body.b.size(), 3);
    if (g_zstd.iserr(r)) { abc_set_error(c, "zstd compress failed"); return ABC_ERR_STATE; }
    packed.resize(r);
    payload = &packed;
  } else if (compr != ABC_SEAL_COMPR_NONE) {
    abc_set_error(c, "unknown compr_mode"); return ABC_ERR_PARAM;
  }
  const u64 size = 16 + payload->size();
  if (len) *len = (size_t)size;
  if (!buf || cap < size) { abc_set_error(c, "output buffer too small (needed size returned)"); return ABC_ERR_PARAM; }
  const unsigned char hdr[8] = {0x5E, 0xA1, 0x10, 3, 6, (unsigned char)compr, 0, 0};
  memcpy(buf, hdr, 8); memcpy(buf + 8, &size, 8);
  memcpy(buf + 16, payload->data(), payload->size());
  return ABC_OK;
}
// stream -> body bytes of the outermost record
abc_status open_record(abc_ctx *c, const uint8_t *bytes, size_t len, std::vector<unsigned char> &body) {
  auto bad = [&](const char *m) { abc_set_error(c, m); return ABC_ERR_PARAM; };
  if (!bytes || len < 16) return bad("SEAL stream truncated");
  if (bytes[0] != 0x5E || bytes[1] != 0xA1 || bytes[2] != 0x10) return bad("not a SEAL record (magic / header size)");
  if (bytes[3] != 3) return bad("SEAL record written by an incompatible major version");
  u64 size; memcpy(&size, bytes + 8, 8);
  if (size < 16 || size > len) return bad("SEAL record size field out of range");
  const unsigned compr = bytes[5];
  std::string err;
  if (compr == ABC_SEAL_COMPR_NONE) body.assign(bytes + 16, bytes + size);
  else if (compr == ABC_SEAL_COMPR_ZLIB) { if (!inflate_all(bytes + 16, size - 16, body, err, max_record_bytes(c))) return bad(err.c_str()); }
  else if (compr == ABC_SEAL_COMPR_ZSTD) { if (!unzstd_all(bytes + 16, size - 16, body, err, max_record_bytes(c))) return bad(err.c_str()); }
  else return bad("unknown compr_mode in SEAL record");
  return ABC_OK;
}

// ---------------------------------------------------------------- parameters of a context
struct Parms { u64 N, t; std::vector<u64> q; };  // q = key-level primes (special prime last)
Parms parms_of(const abc_ctx *c) {
  Parms p;
  p.N = abc_poly_degree(c); p.t = abc_plain_modulus(c);
  p.q.resize(abc_n_primes(c));
  abc_get_primes(c, (uint64_t *)p.q.data());
  return p;
}
// EncryptionParameters::compute_parms_id: hash of [scheme (bfv = 1), N, q_0.., t]; data level drops the special prime
void parms_id(const Parms &p, bool key_level, u64 out[4]) {
  std::vector<u64> d = {1, p.N};
  for (size_t i = 0; i + (key_level ? 0 : 1) < p.q.size(); ++i) d.push_back(p.q[i]);
  d.push_back(p.t);
  blake2b_256(d.data(), d.size() * 8, out);
}

// DynArray<u64>::save_members inside its own record
void put_dynarray(Writer &w, const u64 *data, size_t count) {
  w.record([&](Writer &r) { r.u64v(count); r.raw(data, count * 8); });
}
bool get_dynarray(Reader &r, std::vector<u64> &out, size_t max_count) {
  const u64 size = r.header_plain();
  if (!r.err.empty()) return false;
  const u64 count = r.u64v();
  if (!r.err.empty()) return false;
  if (count > max_count || size != 24 + 8 * count) { r.err = "SEAL DynArray size mismatch"; return false; }
  out.resize(count);
  return count == 0 || r.raw(out.data(), count * 8);
}
// Ciphertext::save_members (no seed): parms_id, is_ntt_form, size, N, coeff_modulus_size, scale, DynArray
void put_ciphertext_body(Writer &w, const u64 id[4], bool ntt, u64 size, u64 N, u64 cms, const u64 *data) {
  w.raw(id, 32); w.u8(ntt ? 1 : 0); w.u64v(size); w.u64v(N); w.u64v(cms); w.f64(1.0);
  put_dynarray(w, data, size * N * cms);
}
bool get_ciphertext_body(Reader &r, const u64 want_id[4], bool want_ntt, u64 want_size, u64 N, const u64 *q, u64 cms,
                         std::vector<u64> &data, const char *what) {
  u64 id[4];
  if (!r.raw(id, 32)) return false;
  const bool ntt = r.u8() != 0;
  const u64 size = r.u64v(), n = r.u64v(), c = r.u64v();
  r.f64();
  if (!r.err.empty()) return false;
  if (memcmp(id, want_id, 32)) { r.err = std::string(what) + ": parms_id differs from this context's encryption parameters"; return false; }
  if (ntt != want_ntt) { r.err = std::string(what) + ": unexpected NTT-form flag"; return false; }
  if (n != N || c != cms) { r.err = std::string(what) + ": poly_modulus_degree / coeff_modulus_size mismatch"; return false; }
  if (size != want_size) { r.err = std::string(what) + ": only size-2 ciphertexts are supported"; return false; }
  if (!get_dynarray(r, data, size * N * cms)) return false;
  if (data.size() == N * cms) { r.err = std::string(what) + ": seeded (symmetric-key) ciphertexts are not supported"; return false; }
  if (data.size() != size * N * cms) { r.err = std::string(what) + ": data length mismatch"; return false; }
  for (u64 pidx = 0; pidx < size; ++pidx)
    for (u64 i = 0; i < cms; ++i) {
      const u64 *row = &data[(pidx * cms + i) * N];
      for (u64 j = 0; j < N; ++j)
        if (row[j] >= q[i]) { r.err = std::string(what) + ": coefficient out of range for its modulus"; return false; }
    }
  return true;
}
}  // namespace

extern "C" {

abc_status abc_seal_parms_id(abc_ctx *c, int key_level, uint64_t out[4]) {
  u64 id[4];
  parms_id(parms_of(c), key_level != 0, id);
  memcpy(out, id, 32);
  return ABC_OK;
}

// EncryptionParameters::save_members: scheme u8, N u64, coeff_modulus_size u64, Modulus records, plain Modulus record
abc_status abc_seal_params_save(abc_ctx *c, int compr, uint8_t *buf, size_t cap, size_t *len) {
  const Parms p = parms_of(c);
  Writer w;
  w.u8(1); w.u64v(p.N); w.u64v(p.q.size());
  for (u64 q : p.q) w.record([&](Writer &r) { r.u64v(q); });
  w.record([&](Writer &r) { r.u64v(p.t); });
  return emit(c, w, compr, buf, cap, len);
}
abc_status abc_seal_params_parse(const uint8_t *bytes, size_t len, abc_params *out, uint64_t *primes_out, size_t primes_cap) {
  std::vector<unsigned char> body;
  if (open_record(nullptr, bytes, len, body) != ABC_OK) return ABC_ERR_PARAM;
  Reader r(body.data(), body.size());
  const unsigned scheme = r.u8();
  const u64 N = r.u64v(), k = r.u64v();
  if (!r.err.empty() || scheme != 1 || k == 0 || k > primes_cap || N > (1u << 20)) {
    abc_set_error(nullptr, scheme != 1 ? "EncryptionParameters: only scheme_type::bfv is supported" : "EncryptionParameters: malformed");
    return ABC_ERR_PARAM;
  }
  for (u64 i = 0; i <= k; ++i) {
    r.header_plain();
    const u64 v = r.u64v();
    if (!r.err.empty()) { abc_set_error(nullptr, r.err.c_str()); return ABC_ERR_PARAM; }
    if (i < k) primes_out[i] = v; else out->plain_modulus = v;
  }
  out->poly_degree = (uint32_t)N; out->n_primes = (uint32_t)k; out->primes = primes_out;
  return ABC_OK;
}

abc_status abc_seal_ct_save(abc_ctx *c, const abc_ct *ct, uint32_t instance, int compr, uint8_t *buf, size_t cap, size_t *len) {
  const Parms p = parms_of(c);
  const u64 L = p.q.size() - 1;
  std::vector<u64> data(2 * L * p.N);
  abc_status s = abc_ct_export_instance(c, ct, instance, (uint64_t *)data.data(), data.size());
  if (s != ABC_OK) return s;
  u64 id[4];
  parms_id(p, false, id);
  Writer w;
  put_ciphertext_body(w, id, false, 2, p.N, L, data.data());
  return emit(c, w, compr, buf, cap, len);
}
abc_status abc_seal_ct_load(abc_ctx *c, abc_ct *ct, uint32_t instance, const uint8_t *bytes, size_t len) {
  std::vector<unsigned char> body;
  abc_status s = open_record(c, bytes, len, body);
  if (s != ABC_OK) return s;
  const Parms p = parms_of(c);
  const u64 L = p.q.size() - 1;
  u64 id[4];
  parms_id(p, false, id);
  Reader r(body.data(), body.size());
  std::vector<u64> data;
  if (!get_ciphertext_body(r, id, false, 2, p.N, p.q.data(), L, data, "Ciphertext")) { abc_set_error(c, r.err.c_str()); return ABC_ERR_PARAM; }
  return abc_ct_import_instance(c, ct, instance, (const uint64_t *)data.data(), data.size());
}

// SecretKey = Plaintext record {parms_id, coeff_count, scale, DynArray}; PublicKey = Ciphertext record;
// KSwitchKeys {parms_id, dim1, per index: dim2, PublicKey records}; GaloisKeys index = (elt - 1) / 2
abc_status abc_seal_key_save(abc_ctx *c, int kind, int compr, uint8_t *buf, size_t cap, size_t *len) {
  const Parms p = parms_of(c);
  const u64 k = p.q.size(), L = k - 1, N = p.N;
  u64 id[4];
  parms_id(p, true, id);
  Writer w;
  abc_status s;
  if (kind == ABC_KEY_SECRET) {
    std::vector<u64> d(k * N);
    if ((s = abc_key_export(c, kind, 0, (uint64_t *)d.data(), d.size())) != ABC_OK) return s;
    w.record([&](Writer &r) { r.raw(id, 32); r.u64v(k * N); r.f64(1.0); put_dynarray(r, d.data(), d.size()); });
  } else if (kind == ABC_KEY_PUBLIC) {
    std::vector<u64> d(2 * k * N);
    if ((s = abc_key_export(c, kind, 0, (uint64_t *)d.data(), d.size())) != ABC_OK) return s;
    w.record([&](Writer &r) { put_ciphertext_body(r, id, true, 2, N, k, d.data()); });
  } else if (kind == ABC_KEY_RELIN || kind == ABC_KEY_GALOIS) {
    std::vector<uint32_t> elts;
    if (kind == ABC_KEY_GALOIS) {
      size_t n = 0;
      abc_galois_elts(c, nullptr, 0, &n);
      elts.resize(n);
      if ((s = abc_galois_elts(c, elts.data(), n, &n)) != ABC_OK) return s;
    }
    // KeyGenerator::create_galois_keys sizes the outer vector to N (one slot per odd element), RelinKeys to 1
    const u64 dim1 = kind == ABC_KEY_GALOIS ? N : 1;
    std::vector<u64> d(L * 2 * k * N);
    w.raw(id, 32); w.u64v(dim1);
    for (u64 idx = 0; idx < dim1; ++idx) {
      uint32_t elt = 0;
      bool have = kind == ABC_KEY_RELIN;
      for (uint32_t e : elts) if (((u64)(e - 1) >> 1) == idx) { elt = e; have = true; }
      if (!have) { w.u64v(0); continue; }
      if ((s = abc_key_export(c, kind, elt, (uint64_t *)d.data(), d.size())) != ABC_OK) return s;
      w.u64v(L);
      for (u64 J = 0; J < L; ++J) {
        const u64 *kj = &d[J * 2 * k * N];
        w.record([&](Writer &pk) { pk.record([&](Writer &r) { put_ciphertext_body(r, id, true, 2, N, k, kj); }); });
      }
    }
  } else {
    abc_set_error(c, "invalid key kind"); return ABC_ERR_PARAM;
  }
  // SecretKey::save / PublicKey::save wrap the inner Plaintext / Ciphertext record: their outer body IS that record
  return emit(c, w, compr, buf, cap, len);
}

abc_status abc_seal_key_load(abc_ctx *c, int kind, const uint8_t *bytes, size_t len) {
  std::vector<unsigned char> body;
  abc_status s = open_record(c, bytes, len, body);
  if (s != ABC_OK) return s;
  const Parms p = parms_of(c);
  const u64 k = p.q.size(), L = k - 1, N = p.N;
  u64 id[4];
  parms_id(p, true, id);
  Reader r(body.data(), body.size());
  auto bad = [&]() { abc_set_error(c, r.err.c_str()); return ABC_ERR_PARAM; };
  std::vector<u64> d;
  if (kind == ABC_KEY_SECRET) {
    r.header_plain();
    u64 got[4];
    r.raw(got, 32);
    const u64 cc = r.u64v();
    r.f64();
    if (!r.err.empty()) return bad();
    if (memcmp(got, id, 32)) { r.err = "SecretKey: parms_id differs from this context's key-level parameters"; return bad(); }
    if (cc != k * N || !get_dynarray(r, d, k * N) || d.size() != k * N) { if (r.err.empty()) r.err = "SecretKey: size mismatch"; return bad(); }
    for (u64 i = 0; i < k; ++i) for (u64 j = 0; j < N; ++j) if (d[i * N + j] >= p.q[i]) { r.err = "SecretKey: coefficient out of range"; return bad(); }
    return abc_key_import(c, kind, 0, (const uint64_t *)d.data(), d.size());
  }
  if (kind == ABC_KEY_PUBLIC) {
    r.header_plain();
    if (!get_ciphertext_body(r, id, true, 2, N, p.q.data(), k, d, "PublicKey")) return bad();
    return abc_key_import(c, kind, 0, (const uint64_t *)d.data(), d.size());
  }
  if (kind != ABC_KEY_RELIN && kind != ABC_KEY_GALOIS) { abc_set_error(c, "invalid key kind"); return ABC_ERR_PARAM; }
  u64 got[4];
  r.raw(got, 32);
  const u64 dim1 = r.u64v();
  if (!r.err.empty()) return bad();
  if (memcmp(got, id, 32)) { r.err = "KSwitchKeys: parms_id differs from this context's key-level parameters"; return bad(); }
  if (dim1 > N || (kind == ABC_KEY_RELIN && dim1 != 1)) { r.err = "KSwitchKeys: unexpected key count (only relinearisation of size-3 ciphertexts)"; return bad(); }
  std::vector<u64> all(L * 2 * k * N);
  size_t loaded = 0;
  for (u64 idx = 0; idx < dim1; ++idx) {
    const u64 dim2 = r.u64v();
    if (!r.err.empty()) return bad();
    if (dim2 == 0) continue;
    if (dim2 != L) { r.err = "KSwitchKeys: decomposition count differs from the number of data primes"; return bad(); }
    for (u64 J = 0; J < L; ++J) {
      r.header_plain();   // PublicKey record
      r.header_plain();   // its Ciphertext record
      if (!get_ciphertext_body(r, id, true, 2, N, p.q.data(), k, d, "KSwitchKeys entry")) return bad();
      memcpy(&all[J * 2 * k * N], d.data(), d.size() * 8);
    }
    const uint32_t elt = kind == ABC_KEY_GALOIS ? (uint32_t)(2 * idx + 1) : 0;
    if ((s = abc_key_import(c, kind, elt, (const uint64_t *)all.data(), all.size())) != ABC_OK) return s;
    ++loaded;
  }
  if (!loaded) { r.err = "KSwitchKeys: stream holds no keys"; return bad(); }
  return ABC_OK;
}

}  // extern "C"
