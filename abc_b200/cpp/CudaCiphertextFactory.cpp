// CudaCiphertextFactory — the factory half of the drop-in for SealCiphertextFactory
// (/root/reference/src/runtime/SealCiphertextFactory.cpp).  Everything that touches ciphertext data calls the C ABI of
// include/abc_b200.h; there is no arithmetic in this file.
// Similarity note: createCiphertext(std::unique_ptr<AbstractValue>&&) and getString are observable-output glue — the
// Cleartext<int>-only rule with its error text, and the "[ a,  b ]" string format — and therefore follow
// SealCiphertextFactory.cpp:174-202 closely on purpose; a drop-in must produce the same strings and throw at the same
// places.  The rest (batch tables, pinned staging, SEAL streams, createPlaintext handles) has no counterpart there.
#include "CudaCiphertextFactory.h"

#include <cstdlib>
#include <sstream>
#include <stdexcept>

#include "CudaCiphertext.h"
#include "abc_b200.h"
#include "ast_opt/runtime/Cleartext.h"

void CudaCiphertextFactory::check(int status) const {
  if (status != ABC_OK) throw std::runtime_error(abc_last_error(ctx));
}

void CudaCiphertextFactory::checkResult(int status, const abc_ct *result) const {
  check(status);
  if (!st->throwOnTransparent) return;
  std::vector<int32_t> flags(abc_batch(ctx));
  check(abc_is_transparent(ctx, result, flags.data()));
  for (int32_t f : flags)
    if (f) throw std::logic_error("result ciphertext is transparent");  // SEAL 3.6.5 Evaluator, same type and text
}

void CudaCiphertextFactory::setup(int device, unsigned int batch, uint64_t seed, const uint8_t *rngKey32) {
  // SealCiphertextFactory::setupSealContext (src/runtime/SealCiphertextFactory.cpp:72-100): BFVDefault(N),
  // Batching(N, 20), then secret/public/Galois/relin keys.  n_primes = 0 and plain_modulus = 0 select the
  // same SEAL defaults inside the library.
  abc_params p{};
  p.poly_degree = ciphertextSlotSize;
  p.device = device;
  p.batch = batch;
  p.seed = seed;
  if (abc_ctx_create(&p, &ctx) != ABC_OK) {
    throw std::runtime_error(std::string("CudaCiphertextFactory: ") + abc_last_error(nullptr));
  }
  st->ctx = ctx;
  if (rngKey32) check(abc_set_rng_key(ctx, rngKey32));
  check(abc_keygen(ctx));
  if (const char *e = std::getenv("ABC_THROW_ON_TRANSPARENT")) st->throwOnTransparent = std::atoi(e) != 0;
}

CudaCiphertextFactory::CudaCiphertextFactory(unsigned int numElementsPerCiphertextSlot, int device, unsigned int batch,
                                             const std::array<uint8_t, 32> &rngKey)
    : ciphertextSlotSize(numElementsPerCiphertextSlot) {
  setup(device, batch, 0, rngKey.data());
}

// The reference's factory draws its keys from SEAL's randomly seeded PRNG; seed 0 asks the library for the same
// behaviour (key seed and encryption salt from the OS generator).  A fixed seed is for tests and for sharing keys
// between the per-GPU factories of one job (extended constructor).
CudaCiphertextFactory::CudaCiphertextFactory() { setup(0, 1, 0); }

CudaCiphertextFactory::CudaCiphertextFactory(unsigned int numElementsPerCiphertextSlot)
    : ciphertextSlotSize(numElementsPerCiphertextSlot) {
  setup(0, 1, 0);
}

CudaCiphertextFactory::CudaCiphertextFactory(unsigned int numElementsPerCiphertextSlot, int device,
                                             unsigned int batch, uint64_t seed)
    : ciphertextSlotSize(numElementsPerCiphertextSlot) {
  setup(device, batch, seed);
}

CudaCiphertextFactory::State::~State() {
  for (auto &t : batchTables) abc_host_unregister(t.data());
  abc_host_free(pinnedOut);
  abc_ctx_destroy(ctx);   // completed by the last live ciphertext handle, if any (abc_b200.h)
}

void CudaCiphertextFactory::releaseBatchTables() const {
  for (auto &t : st->batchTables) abc_host_unregister(t.data());
  st->batchTables.clear();
  st->batchTableWidth.clear();
  st->nextBatchTable = 0;
}

std::unique_ptr<AbstractCiphertext> CudaCiphertextFactory::loadCiphertext(const std::vector<uint8_t> &sealStream,
                                                                          unsigned int instance) const {
  auto result = std::make_unique<CudaCiphertext>(*this);
  check(abc_seal_ct_load(ctx, result->getHandle(), instance, sealStream.data(), sealStream.size()));
  return result;
}
std::vector<uint8_t> CudaCiphertextFactory::saveCiphertext(const AbstractCiphertext &abstractCiphertext,
                                                           unsigned int instance, int compr) const {
  auto c = dynamic_cast<const CudaCiphertext *>(&abstractCiphertext);
  if (!c) throw std::runtime_error("Cast of AbstractCiphertext to CudaCiphertext failed!");
  size_t len = 0;
  abc_seal_ct_save(ctx, c->getHandle(), instance, compr, nullptr, 0, &len);  // size query
  std::vector<uint8_t> out(len);
  check(abc_seal_ct_save(ctx, c->getHandle(), instance, compr, out.data(), out.size(), &len));
  out.resize(len);
  return out;
}
void CudaCiphertextFactory::loadKey(int kind, const std::vector<uint8_t> &sealStream) const {
  check(abc_seal_key_load(ctx, kind, sealStream.data(), sealStream.size()));
}
std::vector<uint8_t> CudaCiphertextFactory::saveKey(int kind, int compr) const {
  size_t len = 0;
  abc_seal_key_save(ctx, kind, compr, nullptr, 0, &len);
  std::vector<uint8_t> out(len);
  check(abc_seal_key_save(ctx, kind, compr, out.data(), out.size(), &len));
  out.resize(len);
  return out;
}

CudaPlaintext::~CudaPlaintext() { abc_pt_free(handle); }

std::unique_ptr<CudaPlaintext> CudaCiphertextFactory::createPlaintext(const std::vector<int64_t> &value) const {
  abc_pt *h = nullptr;
  check(abc_pt_encode(ctx, value.data(), value.size(), /*broadcast=*/1, &h));   // empty / oversize vectors: same errors as createCiphertext
  return std::make_unique<CudaPlaintext>(h);
}
std::unique_ptr<CudaPlaintext> CudaCiphertextFactory::createPlaintext(const std::vector<int> &value) const {
  return createPlaintext(std::vector<int64_t>(value.begin(), value.end()));
}
std::unique_ptr<CudaPlaintext> CudaCiphertextFactory::createPlaintext(int64_t value) const {
  return createPlaintext(std::vector<int64_t>{value});
}
std::unique_ptr<AbstractCiphertext> CudaCiphertextFactory::encryptPlaintext(const CudaPlaintext &plaintext) const {
  abc_ct *h = nullptr;
  check(abc_encrypt_pt(ctx, plaintext.handle, &h));
  return std::make_unique<CudaCiphertext>(*this, h);
}

unsigned int CudaCiphertextFactory::getCiphertextSlotSize() const { return ciphertextSlotSize; }
unsigned int CudaCiphertextFactory::getBatchSize() const { return abc_batch(ctx); }
void CudaCiphertextFactory::synchronize() const { check(abc_sync(ctx)); }
uint64_t CudaCiphertextFactory::launchCount() const { return abc_launch_count(ctx); }

std::unique_ptr<AbstractCiphertext> CudaCiphertextFactory::createCiphertext(const std::vector<int64_t> &data) const {
  // expandVector + BatchEncoder::encode + Encryptor::encrypt, all on the device; an empty vector is an
  // error here (the reference calls .back() on it, SealCiphertextFactory.cpp:112).
  abc_ct *h = nullptr;
  check(abc_encode_encrypt(ctx, data.data(), data.size(), /*broadcast=*/1, &h));
  return std::make_unique<CudaCiphertext>(*this, h);
}

std::unique_ptr<AbstractCiphertext> CudaCiphertextFactory::createCiphertext(const std::vector<int> &data) const {
  std::vector<int64_t> ciphertextData(data.begin(), data.end());
  return createCiphertext(ciphertextData);
}

std::unique_ptr<AbstractCiphertext> CudaCiphertextFactory::createCiphertext(int64_t data) const {
  std::vector<int64_t> values = {data};
  return createCiphertext(values);
}

std::unique_ptr<AbstractCiphertext> CudaCiphertextFactory::createCiphertext(
    std::unique_ptr<AbstractValue> &&abstractValue) const {
  if (auto castedCleartext = dynamic_cast<Cleartext<int> *>(abstractValue.get())) {
    auto castedCleartextData = castedCleartext->getData();
    if (st->nextBatchTable < st->batchTables.size()) {   // lock-step batch: this declaration's values for every instance
      // the table decides the values AND their count: the literal in the program text is a placeholder (a one-element
      // literal keeps the interpreter from evaluating thousands of LiteralInt nodes per declaration: 28 ms each at n = 4096)
      const size_t d = st->nextBatchTable++;
      return createCiphertextBatch(st->batchTables[d], st->batchTableWidth[d]);
    }
    std::vector<int64_t> data(castedCleartextData.begin(), castedCleartextData.end());
    return createCiphertext(data);
  }
  throw std::runtime_error("Cannot create ciphertext from any other than a Cleartext<int> as used ciphertext factory "
                           "(CudaCiphertextFactory) uses BFV that only supports integers.");
}

void CudaCiphertextFactory::setBatchInputs(std::vector<std::vector<int64_t>> tables) const {
  const size_t B = getBatchSize();
  std::vector<size_t> widths;
  for (const auto &t : tables) {
    if (t.empty() || t.size() % B) throw std::runtime_error("setBatchInputs: every table needs batch * n values");
    widths.push_back(t.size() / B);
  }
  releaseBatchTables();
  st->batchTableWidth = std::move(widths);
  st->batchTables = std::move(tables);
  for (auto &t : st->batchTables) check(abc_host_register(ctx, t.data(), t.size() * sizeof(int64_t)));   // H2D straight from the tables
  st->nextBatchTable = 0;
}

const int64_t *CudaCiphertextFactory::decryptCiphertextBatchPinnedAsync(AbstractCiphertext &abstractCiphertext) const {
  auto c = dynamic_cast<CudaCiphertext *>(&abstractCiphertext);
  if (!c) throw std::runtime_error("Cast of AbstractCiphertext to CudaCiphertext failed!");
  const size_t words = static_cast<size_t>(getBatchSize()) * ciphertextSlotSize;
  if (!st->pinnedOut) {
    void *p = nullptr;
    check(abc_host_alloc(ctx, 2 * words * sizeof(int64_t), &p));
    st->pinnedOut = static_cast<int64_t *>(p);
  }
  int64_t *dst = st->pinnedOut + (st->pinnedNext++ & 1u) * words;
  check(abc_decrypt_decode_async(ctx, c->getHandle(), dst));
  return dst;
}
void CudaCiphertextFactory::waitDecryptions() const { check(abc_decrypt_wait(ctx)); }
const int64_t *CudaCiphertextFactory::decryptCiphertextBatchPinned(AbstractCiphertext &abstractCiphertext) const {
  const int64_t *r = decryptCiphertextBatchPinnedAsync(abstractCiphertext);
  waitDecryptions();
  return r;
}

std::unique_ptr<AbstractCiphertext> CudaCiphertextFactory::createCiphertextBatch(const std::vector<int64_t> &data,
                                                                                 size_t n) const {
  if (n == 0 || data.size() != n * getBatchSize()) throw std::runtime_error("createCiphertextBatch: need batch*n values");
  abc_ct *h = nullptr;
  check(abc_encode_encrypt(ctx, data.data(), n, /*broadcast=*/0, &h));
  return std::make_unique<CudaCiphertext>(*this, h);
}

static CudaCiphertext &castCuda(AbstractCiphertext &abstractCiphertext) {
  if (auto c = dynamic_cast<CudaCiphertext *>(&abstractCiphertext)) return *c;
  throw std::runtime_error("Cast of AbstractCiphertext to CudaCiphertext failed!");
}

void CudaCiphertextFactory::decryptCiphertextBatch(AbstractCiphertext &abstractCiphertext,
                                                   std::vector<int64_t> &out) const {
  auto &ctxt = castCuda(abstractCiphertext);
  out.assign(static_cast<size_t>(getBatchSize()) * ciphertextSlotSize, 0);
  check(abc_decrypt_decode(ctx, ctxt.getHandle(), out.data()));
}

void CudaCiphertextFactory::decryptCiphertext(AbstractCiphertext &abstractCiphertext,
                                              std::vector<int64_t> &ciphertextData) const {
  // Decryptor::decrypt + BatchEncoder::decode: overwrites the vector with exactly N signed values
  // (instance 0 when the factory was built with a batch).
  decryptCiphertextBatch(abstractCiphertext, ciphertextData);
  ciphertextData.resize(ciphertextSlotSize);
}

std::vector<uint64_t> CudaCiphertextFactory::exportCoefficients(const AbstractCiphertext &abstractCiphertext) const {
  auto c = dynamic_cast<const CudaCiphertext *>(&abstractCiphertext);
  if (!c) throw std::runtime_error("Cast of AbstractCiphertext to CudaCiphertext failed!");
  std::vector<uint64_t> out(abc_ct_words(ctx));
  check(abc_ct_export(ctx, c->getHandle(), out.data(), out.size()));
  return out;
}

std::string CudaCiphertextFactory::getString(AbstractCiphertext &abstractCiphertext) const {
  std::vector<int64_t> plainValues;
  decryptCiphertext(abstractCiphertext, plainValues);
  std::stringstream ss;
  ss << "[";
  for (const auto value : plainValues) ss << " " << value << ", ";
  ss.seekp(-1, ss.cur);
  ss << " ]";
  return ss.str();
}
