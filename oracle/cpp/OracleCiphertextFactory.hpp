// OracleCiphertextFactory / OracleCiphertext — TEST INFRASTRUCTURE (CPU baseline arm), never shipped, never on the
// product path.  The CPU oracle (oracle/bfv_oracle.c: SEAL 3.6.5's BFV algorithms restated in C) behind ABC's
// AbstractCiphertextFactory / AbstractCiphertext (include/ast_opt/runtime/AbstractCiphertextFactory.h:19-49,
// AbstractCiphertext.h:27-98), the way SealCiphertextFactory / SealCiphertext put SEAL behind them
// (src/runtime/SealCiphertextFactory.cpp:72-152, src/runtime/SealCiphertext.cpp:52-202) — so that the CPU arm of
// bench.py runs through the reference's own RuntimeVisitor, as BASELINE.md 3.1 states.  SEAL itself is not installable here.
// Several factories may share one obfv_ctx (parameters + keys, read-only after keygen): one factory per host thread.
#pragma once
#include <atomic>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../bfv_oracle.h"
#include "ast_opt/runtime/AbstractCiphertext.h"
#include "ast_opt/runtime/AbstractCiphertextFactory.h"
#include "ast_opt/runtime/Cleartext.h"

class OracleCiphertextFactory;

class OracleCiphertext : public AbstractCiphertext {
 public:
  std::vector<uint64_t> ct;   // [2][L][N] coefficient form, canonical residues (seal::Ciphertext's data for BFV)
  explicit OracleCiphertext(const std::reference_wrapper<const OracleCiphertextFactory> f);
  OracleCiphertext(const OracleCiphertext &other) = default;
  [[nodiscard]] const AbstractCiphertextFactory &getFactory() const override { return factory.get(); }
  [[nodiscard]] const OracleCiphertextFactory &oracleFactory() const;
  obfv_ctx *ctx() const;
  static const OracleCiphertext &cast(const AbstractCiphertext &a) {
    if (auto c = dynamic_cast<const OracleCiphertext *>(&a)) return *c;
    throw std::runtime_error("Cast of AbstractCiphertext to OracleCiphertext failed!");
  }
  std::vector<uint64_t> plainOf(const ICleartext &operand, const char *op) const;

  [[nodiscard]] std::unique_ptr<AbstractCiphertext> multiply(const AbstractCiphertext &operand) const override {
    auto r = std::make_unique<OracleCiphertext>(*this); r->multiplyInplace(operand); return r;
  }
  void multiplyInplace(const AbstractCiphertext &operand) override {   // multiply_inplace + relinearize_inplace (SealCiphertext.cpp:121-124)
    std::vector<uint64_t> p3(ct.size() / 2 * 3);
    obfv_multiply(ctx(), ct.data(), cast(operand).ct.data(), p3.data());
    obfv_relinearize(ctx(), p3.data(), ct.data());
  }
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> add(const AbstractCiphertext &operand) const override {
    auto r = std::make_unique<OracleCiphertext>(*this); r->addInplace(operand); return r;
  }
  void addInplace(const AbstractCiphertext &operand) override { obfv_add(ctx(), ct.data(), cast(operand).ct.data(), ct.data()); }
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> subtract(const AbstractCiphertext &operand) const override {
    auto r = std::make_unique<OracleCiphertext>(*this); r->subtractInplace(operand); return r;
  }
  void subtractInplace(const AbstractCiphertext &operand) override { obfv_sub(ctx(), ct.data(), cast(operand).ct.data(), ct.data()); }
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> rotateRows(int steps) const override {
    auto r = std::make_unique<OracleCiphertext>(*this); r->rotateRowsInplace(steps); return r;
  }
  void rotateRowsInplace(int steps) override {
    std::vector<uint64_t> out(ct.size());
    if (obfv_rotate_rows(ctx(), ct.data(), steps, out.data()) != 0) throw std::runtime_error("step count too large");
    ct.swap(out);
  }
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> multiplyPlain(const ICleartext &operand) const override {
    auto r = std::make_unique<OracleCiphertext>(*this); r->multiplyPlainInplace(operand); return r;
  }
  void multiplyPlainInplace(const ICleartext &operand) override {
    auto p = plainOf(operand, "MULTIPLY");
    std::vector<uint64_t> out(ct.size());
    obfv_multiply_plain(ctx(), ct.data(), p.data(), out.data());
    ct.swap(out);
  }
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> addPlain(const ICleartext &operand) const override {
    auto r = std::make_unique<OracleCiphertext>(*this); r->addPlainInplace(operand); return r;
  }
  void addPlainInplace(const ICleartext &operand) override {
    auto p = plainOf(operand, "ADD");
    obfv_add_plain(ctx(), ct.data(), p.data(), ct.data());
  }
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> subtractPlain(const ICleartext &operand) const override {
    auto r = std::make_unique<OracleCiphertext>(*this); r->subtractPlainInplace(operand); return r;
  }
  void subtractPlainInplace(const ICleartext &operand) override {
    auto p = plainOf(operand, "SUBTRACT");
    obfv_sub_plain(ctx(), ct.data(), p.data(), ct.data());
  }
  std::unique_ptr<AbstractCiphertext> clone() const override { return std::make_unique<OracleCiphertext>(*this); }

  // AbstractValue dispatch (SealCiphertext.cpp:204-239)
  void add_inplace(const AbstractValue &other) override {
    if (auto c = dynamic_cast<const OracleCiphertext *>(&other)) addInplace(*c);
    else if (auto p = dynamic_cast<const ICleartext *>(&other)) addPlainInplace(*p);
    else throw std::runtime_error("Operation ADD only supported for (OracleCiphertext,OracleCiphertext) and (OracleCiphertext, ICleartext).");
  }
  void subtract_inplace(const AbstractValue &other) override {
    if (auto c = dynamic_cast<const OracleCiphertext *>(&other)) subtractInplace(*c);
    else if (auto p = dynamic_cast<const ICleartext *>(&other)) subtractPlainInplace(*p);
    else throw std::runtime_error("Operation SUBTRACT only supported for (OracleCiphertext,OracleCiphertext) and (OracleCiphertext, ICleartext).");
  }
  void multiply_inplace(const AbstractValue &other) override {
    if (auto c = dynamic_cast<const OracleCiphertext *>(&other)) multiplyInplace(*c);
    else if (auto p = dynamic_cast<const ICleartext *>(&other)) multiplyPlainInplace(*p);
    else throw std::runtime_error("Operation MULTIPLY only supported for (OracleCiphertext,OracleCiphertext) and (OracleCiphertext, ICleartext).");
  }
#define ORACLE_UNSUPPORTED(NAME) \
  void NAME(const AbstractValue &) override { throw std::runtime_error("Operation " #NAME " not supported for (OracleCiphertext, ANY)."); }
  ORACLE_UNSUPPORTED(divide_inplace) ORACLE_UNSUPPORTED(modulo_inplace) ORACLE_UNSUPPORTED(logicalAnd_inplace)
  ORACLE_UNSUPPORTED(logicalOr_inplace) ORACLE_UNSUPPORTED(logicalLess_inplace) ORACLE_UNSUPPORTED(logicalLessEqual_inplace)
  ORACLE_UNSUPPORTED(logicalGreater_inplace) ORACLE_UNSUPPORTED(logicalGreaterEqual_inplace) ORACLE_UNSUPPORTED(logicalEqual_inplace)
  ORACLE_UNSUPPORTED(logicalNotEqual_inplace) ORACLE_UNSUPPORTED(bitwiseAnd_inplace) ORACLE_UNSUPPORTED(bitwiseXor_inplace)
  ORACLE_UNSUPPORTED(bitwiseOr_inplace)
#undef ORACLE_UNSUPPORTED
  void logicalNot_inplace() override { throw std::runtime_error("Operation logicalNot_inplace not supported for (OracleCiphertext, ANY)."); }
  void bitwiseNot_inplace() override { throw std::runtime_error("Operation bitwiseNot_inplace not supported for (OracleCiphertext, ANY)."); }
};

class OracleCiphertextFactory : public AbstractCiphertextFactory {
 public:
  std::shared_ptr<obfv_ctx> ctx;                    // shared between the per-thread factories
  std::shared_ptr<std::atomic<uint64_t>> nonce;     // encryption randomness counter, shared with them too
  size_t N;
  // per-declaration inputs (the same hook as CudaCiphertextFactory::setBatchInputs, one instance): the next
  // createCiphertext(unique_ptr<AbstractValue>&&) calls take these values instead of the (placeholder) literal
  mutable std::vector<std::vector<int64_t>> nextInputs;
  mutable size_t nextInput = 0;

  explicit OracleCiphertextFactory(unsigned int numElementsPerCiphertextSlot, uint64_t seed = 4673838) : N(numElementsPerCiphertextSlot) {
    obfv_ctx *c = obfv_create(N, nullptr, 0, 0);    // BFVDefault(N), Batching(N, 20): SealCiphertextFactory.cpp:72-100
    if (!c) throw std::runtime_error("OracleCiphertextFactory: invalid parameters");
    obfv_keygen(c, seed);
    ctx = std::shared_ptr<obfv_ctx>(c, obfv_destroy);
    nonce = std::make_shared<std::atomic<uint64_t>>(0);
  }
  OracleCiphertextFactory(const OracleCiphertextFactory &other) = default;   // shares context, keys and nonce counter

  void setNextInputs(std::vector<std::vector<int64_t>> v) const { nextInputs = std::move(v); nextInput = 0; }

  std::unique_ptr<AbstractCiphertext> createCiphertext(const std::vector<int64_t> &data) const override {
    if (data.empty()) throw std::runtime_error("Cannot encode an empty vector.");
    if (data.size() > N)
      throw std::runtime_error("Cannot encode " + std::to_string(data.size()) + " elements in a ciphertext of size " + std::to_string(N) + ". ");
    std::vector<int64_t> slots(data);
    slots.resize(N, data.back());                   // expandVector: pad with the last value (SealCiphertextFactory.cpp:102-118)
    std::vector<uint64_t> plain(N);
    obfv_encode(ctx.get(), slots.data(), plain.data());
    auto r = std::make_unique<OracleCiphertext>(*this);
    obfv_encrypt(ctx.get(), plain.data(), nonce->fetch_add(1), r->ct.data());
    return r;
  }
  std::unique_ptr<AbstractCiphertext> createCiphertext(const std::vector<int> &data) const override {
    return createCiphertext(std::vector<int64_t>(data.begin(), data.end()));
  }
  std::unique_ptr<AbstractCiphertext> createCiphertext(int64_t data) const override { return createCiphertext(std::vector<int64_t>{data}); }
  std::unique_ptr<AbstractCiphertext> createCiphertext(std::unique_ptr<AbstractValue> &&abstractValue) const override {
    if (auto c = dynamic_cast<Cleartext<int> *>(abstractValue.get())) {
      if (nextInput < nextInputs.size()) return createCiphertext(nextInputs[nextInput++]);
      auto d = c->getData();
      return createCiphertext(std::vector<int64_t>(d.begin(), d.end()));
    }
    throw std::runtime_error("Cannot create ciphertext from any other than a Cleartext<int> (BFV supports integers only).");
  }
  void decryptCiphertext(AbstractCiphertext &abstractCiphertext, std::vector<int64_t> &ciphertextData) const override {
    auto &c = OracleCiphertext::cast(abstractCiphertext);
    std::vector<uint64_t> plain(N);
    obfv_decrypt(ctx.get(), c.ct.data(), 2, plain.data());
    ciphertextData.assign(N, 0);
    obfv_decode(ctx.get(), plain.data(), ciphertextData.data());
  }
  std::string getString(AbstractCiphertext &abstractCiphertext) const override {
    std::vector<int64_t> v;
    decryptCiphertext(abstractCiphertext, v);
    std::stringstream ss;
    ss << "[";
    for (auto x : v) ss << " " << x << ", ";
    ss.seekp(-1, ss.cur);
    ss << " ]";
    return ss.str();
  }
};

inline OracleCiphertext::OracleCiphertext(const std::reference_wrapper<const OracleCiphertextFactory> f)
    : AbstractCiphertext((const std::reference_wrapper<const AbstractCiphertextFactory>)f),
      ct(2 * obfv_L(f.get().ctx.get()) * f.get().N) {}
inline const OracleCiphertextFactory &OracleCiphertext::oracleFactory() const {
  if (auto f = dynamic_cast<const OracleCiphertextFactory *>(&factory.get())) return *f;
  throw std::runtime_error("Cast of AbstractFactory to OracleCiphertextFactory failed.");
}
inline obfv_ctx *OracleCiphertext::ctx() const { return oracleFactory().ctx.get(); }
inline std::vector<uint64_t> OracleCiphertext::plainOf(const ICleartext &operand, const char *op) const {
  auto c = dynamic_cast<const Cleartext<int> *>(&operand);
  if (!c) throw std::runtime_error(std::string(op) + "(Ciphertext,Cleartext) requires a Cleartext<int> as BFV supports integers only.");
  const size_t N = oracleFactory().N;
  auto d = c->getData();
  if (d.empty() || d.size() > N) throw std::runtime_error("Cannot encode the cleartext operand.");
  std::vector<int64_t> slots(d.begin(), d.end());
  slots.resize(N, slots.back());
  std::vector<uint64_t> plain(N);
  obfv_encode(oracleFactory().ctx.get(), slots.data(), plain.data());
  return plain;
}
