// CudaCiphertextFactory — B200 drop-in for ABC's SealCiphertextFactory.
//
// Implements AbstractCiphertextFactory (/root/reference/include/ast_opt/runtime/AbstractCiphertextFactory.h:19-49)
// with the same public surface as SealCiphertextFactory (include/ast_opt/runtime/SealCiphertextFactory.h:59-123):
// same constructor argument (slot count; SEAL's BFVDefault coefficient modulus and Batching(N,20) plain modulus),
// same pad-with-last / oversize-throws behaviour, same decrypt semantics.  RuntimeVisitor drives it unchanged.
// All arithmetic runs in libabc_b200.so (include/abc_b200.h); there is no CPU path in this class.
#ifndef ABC_B200_CPP_CUDACIPHERTEXTFACTORY_H_
#define ABC_B200_CPP_CUDACIPHERTEXTFACTORY_H_

#include <cstdint>
#include <array>
#include <memory>
#include <string>
#include <vector>

#include "ast_opt/runtime/AbstractCiphertextFactory.h"

struct abc_ctx;
struct abc_ct;
struct abc_pt;
class CudaCiphertext;

/// An encoded plaintext on the device: what SealCiphertextFactory::createPlaintext returns as a seal::Plaintext
/// (include/ast_opt/runtime/SealCiphertextFactory.h:95-107).  Owns its device handle.
class CudaPlaintext {
 public:
  abc_pt *handle = nullptr;
  explicit CudaPlaintext(abc_pt *h) : handle(h) {}
  ~CudaPlaintext();
  CudaPlaintext(const CudaPlaintext &) = delete;
  CudaPlaintext &operator=(const CudaPlaintext &) = delete;
};

class CudaCiphertextFactory : public AbstractCiphertextFactory {
 private:
  /// The number of slots (= polynomial degree N) of every ciphertext created by this factory.
  const unsigned int ciphertextSlotSize = 16'384;  // same default as SealCiphertextFactory.h:13
  /// Device context (parameters, NTT tables, keys, stream) and the factory's host-side staging, shared between copies of
  /// the factory: like SealCiphertextFactory, this class is copy-constructible — a copy refers to the same keys on the
  /// same device.  All factory virtuals are const, so the mutable state lives behind this pointer.
  struct State {
    abc_ctx *ctx = nullptr;
    // lock-step batch driver: per-instance values of the next `secret` declarations, in declaration order
    std::vector<std::vector<int64_t>> batchTables;   // [declaration][batch * n], instance-major
    std::vector<size_t> batchTableWidth;             // n of each table
    size_t nextBatchTable = 0;
    int64_t *pinnedOut = nullptr;                    // page-locked 2 x batch * N slots: decryptCiphertextBatchPinned
    unsigned pinnedNext = 0;
    bool throwOnTransparent = false;                 // setThrowOnTransparent
    ~State();
  };
  std::shared_ptr<State> st = std::make_shared<State>();
  abc_ctx *ctx = nullptr;                            // = st->ctx
  void releaseBatchTables() const;

  void setup(int device, unsigned int batch, uint64_t seed, const uint8_t *rngKey32 = nullptr);

 public:
  CudaCiphertextFactory();
  /// \param numElementsPerCiphertextSlot slot count N (4096, 8192, 16384 use SEAL's default parameter sets).
  explicit CudaCiphertextFactory(unsigned int numElementsPerCiphertextSlot);
  /// Extended constructor: device ordinal, instances per handle (lock-step batch), sampler seed.
  CudaCiphertextFactory(unsigned int numElementsPerCiphertextSlot, int device, unsigned int batch, uint64_t seed);
  /// The same with the sampler's 256-bit ChaCha20 key supplied by the caller (32 bytes of its own entropy): the way to give
  /// the per-GPU factories of one job the same keys without a guessable 64-bit seed.
  CudaCiphertextFactory(unsigned int numElementsPerCiphertextSlot, int device, unsigned int batch,
                        const std::array<uint8_t, 32> &rngKey);
  ~CudaCiphertextFactory() = default;
  /// A copy shares the device context (keys, stream) with the original; the context goes when the last copy — and the last
  /// ciphertext created through it — is gone.  One host thread at a time per context (copies included).
  CudaCiphertextFactory(const CudaCiphertextFactory &) = default;

  std::unique_ptr<AbstractCiphertext> createCiphertext(const std::vector<int64_t> &data) const override;
  std::unique_ptr<AbstractCiphertext> createCiphertext(const std::vector<int> &data) const override;
  std::unique_ptr<AbstractCiphertext> createCiphertext(int64_t data) const override;
  std::unique_ptr<AbstractCiphertext> createCiphertext(std::unique_ptr<AbstractValue> &&cleartext) const override;
  void decryptCiphertext(AbstractCiphertext &abstractCiphertext, std::vector<int64_t> &ciphertextData) const override;
  std::string getString(AbstractCiphertext &abstractCiphertext) const override;

  /// SealCiphertextFactory::createPlaintext x 3 (SealCiphertextFactory.h:95-107): pad-with-last + BatchEncoder::encode on
  /// the device.  The result can be encrypted later (encryptPlaintext) — the two halves of createCiphertext.
  std::unique_ptr<CudaPlaintext> createPlaintext(const std::vector<int> &value) const;
  std::unique_ptr<CudaPlaintext> createPlaintext(const std::vector<int64_t> &value) const;
  std::unique_ptr<CudaPlaintext> createPlaintext(int64_t value) const;
  std::unique_ptr<AbstractCiphertext> encryptPlaintext(const CudaPlaintext &plaintext) const;

  /// Gets the number of slots of a ciphertext (SealCiphertextFactory::getCiphertextSlotSize).
  [[nodiscard]] unsigned int getCiphertextSlotSize() const;
  /// Instances carried by every ciphertext handle (1 unless constructed with a batch).
  [[nodiscard]] unsigned int getBatchSize() const;
  /// Lock-step batch driver: ONE interpreter walk (RuntimeVisitor unchanged) drives `batch` independent encrypted
  /// programs.  Register, in the order the walk will declare them, the values of every `secret` input for all
  /// instances (table d holds batch * n values, instance-major).  The next createCiphertext(unique_ptr<AbstractValue>&&)
  /// calls — the ones RuntimeVisitor makes for `secret` declarations (src/runtime/RuntimeVisitor.cpp:413) — consume the
  /// tables instead of the literal in the program text (a placeholder: the table decides values and count); calls after the
  /// last table broadcast their literal to every instance as before.
  void setBatchInputs(std::vector<std::vector<int64_t>> tables) const;
  /// The next interpreter walk starts again at the first registered table (same inputs, e.g. a timed repetition).
  void rewindBatchInputs() const { st->nextBatchTable = 0; }
  /// Batched variants: data holds batch*n slot values (instance-major); out gets batch*N values.
  std::unique_ptr<AbstractCiphertext> createCiphertextBatch(const std::vector<int64_t> &data, size_t n) const;
  void decryptCiphertextBatch(AbstractCiphertext &abstractCiphertext, std::vector<int64_t> &out) const;
  /// The same into page-locked staging owned by the factory (no pageable copy of batch * N * 8 bytes): the pointer is
  /// valid until the next call.
  const int64_t *decryptCiphertextBatchPinned(AbstractCiphertext &abstractCiphertext) const;
  /// Enqueued variant: returns where the slots WILL be once waitDecryptions() (or synchronize()) has returned; two staging
  /// buffers alternate, so the D2H copy of one walk's result runs under the next walk's kernels.
  const int64_t *decryptCiphertextBatchPinnedAsync(AbstractCiphertext &abstractCiphertext) const;
  void waitDecryptions() const;
  /// Raw coefficients [batch][2][L][N] of a ciphertext (the bit-exactness probe).
  std::vector<uint64_t> exportCoefficients(const AbstractCiphertext &abstractCiphertext) const;
  /// Microsoft SEAL 3.6 binary streams (seal::Ciphertext::save/load, SecretKey/PublicKey/RelinKeys/GaloisKeys::save/load):
  /// the way artefacts of a SEAL-backed SealCiphertextFactory move to this backend and back.  compr: 0 none, 1 zlib,
  /// 2 zstd.  `instance` selects one ciphertext of a batched handle.  kind: ABC_KEY_SECRET .. ABC_KEY_GALOIS.
  std::unique_ptr<AbstractCiphertext> loadCiphertext(const std::vector<uint8_t> &sealStream, unsigned int instance = 0) const;
  std::vector<uint8_t> saveCiphertext(const AbstractCiphertext &abstractCiphertext, unsigned int instance = 0, int compr = 0) const;
  void loadKey(int kind, const std::vector<uint8_t> &sealStream) const;
  std::vector<uint8_t> saveKey(int kind, int compr = 0) const;
  /// Blocks until all enqueued work of this factory has finished.
  void synchronize() const;
  /// Kernels launched so far.
  [[nodiscard]] uint64_t launchCount() const;

  [[nodiscard]] abc_ctx *context() const { return ctx; }
  /// Throws std::runtime_error(abc_last_error) when status != 0 (the reference's error convention).
  void check(int status) const;
  /// check(status), then — if setThrowOnTransparent(true) — SEAL's verdict on the result of an evaluator op.
  void checkResult(int status, const abc_ct *result) const;
  /// Mirror SEAL's SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT (its default build option): every evaluator op whose result has
  /// an all-zero c1 (`x --- x`, `x *** 0`) throws std::logic_error("result ciphertext is transparent").  Off by default:
  /// the test is a device reduction and a host synchronisation after EVERY op, which serialises the stream the backend
  /// otherwise keeps full.  Also switched on by the environment variable ABC_THROW_ON_TRANSPARENT=1.  With a lock-step
  /// batch the op throws if any instance's result is transparent.
  void setThrowOnTransparent(bool on) const { st->throwOnTransparent = on; }
};

#endif  // ABC_B200_CPP_CUDACIPHERTEXTFACTORY_H_
