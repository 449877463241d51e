// CudaCiphertext — the ciphertext half of the drop-in for SealCiphertext (/root/reference/src/runtime/SealCiphertext.cpp):
// the same virtuals with the same dispatch shape and message texts (a subclass of the same abstract interface has to look
// like its sibling), every body re-expressed as C-ABI calls (include/abc_b200.h) instead of seal::Evaluator calls.
#include "CudaCiphertext.h"

#include <stdexcept>
#include <vector>

#include "abc_b200.h"
#include "ast_opt/runtime/Cleartext.h"

namespace {
const CudaCiphertext &cast(const AbstractCiphertext &abstractCiphertext) {
  if (auto c = dynamic_cast<const CudaCiphertext *>(&abstractCiphertext)) return *c;
  throw std::runtime_error("Cast of AbstractCiphertext to CudaCiphertext failed!");
}
std::vector<int64_t> widen(const std::vector<int> &v) { return std::vector<int64_t>(v.begin(), v.end()); }
const Cleartext<int> &intCleartext(const ICleartext &operand, const char *op) {
  if (auto c = dynamic_cast<const Cleartext<int> *>(&operand)) return *c;
  throw std::runtime_error(std::string(op) + "(Ciphertext,Cleartext) requires a Cleartext<int> as BFV supports integers only.");
}
}  // namespace

CudaCiphertext::CudaCiphertext(const std::reference_wrapper<const CudaCiphertextFactory> cudaFactory)
    : AbstractCiphertext((const std::reference_wrapper<const AbstractCiphertextFactory>)cudaFactory) {
  getFactory().check(abc_ct_alloc(getFactory().context(), &handle));
}

CudaCiphertext::CudaCiphertext(const std::reference_wrapper<const CudaCiphertextFactory> cudaFactory, abc_ct *owned)
    : AbstractCiphertext((const std::reference_wrapper<const AbstractCiphertextFactory>)cudaFactory), handle(owned) {}

CudaCiphertext::~CudaCiphertext() { abc_ct_free(handle); }

CudaCiphertext::CudaCiphertext(const CudaCiphertext &other) : AbstractCiphertext(other.factory) {
  getFactory().check(abc_ct_clone(getFactory().context(), other.handle, &handle));
}

CudaCiphertext::CudaCiphertext(CudaCiphertext &&other) noexcept : AbstractCiphertext(other.factory), handle(other.handle) {
  other.handle = nullptr;
}

CudaCiphertext &CudaCiphertext::operator=(const CudaCiphertext &other) { return *this = CudaCiphertext(other); }

CudaCiphertext &CudaCiphertext::operator=(CudaCiphertext &&other) {
  if (&other == this) return *this;
  if (&factory.get() != &(other.factory.get())) {
    throw std::runtime_error("Cannot move Ciphertext from factory A into Ciphertext created by Factory B.");
  }
  abc_ct_free(handle);
  handle = other.handle;
  other.handle = nullptr;
  return *this;
}

const CudaCiphertextFactory &CudaCiphertext::getFactory() const {
  if (auto cudaFactory = dynamic_cast<const CudaCiphertextFactory *>(&factory.get())) return *cudaFactory;
  throw std::runtime_error("Cast of AbstractFactory to CudaFactory failed. CudaCiphertext is probably invalid.");
}

int CudaCiphertext::noiseBits() const {
  std::vector<int32_t> bits(abc_batch(getFactory().context()));
  getFactory().check(abc_noise_budget(getFactory().context(), handle, bits.data()));
  int m = bits[0];
  for (int b : bits) m = b < m ? b : m;
  return m;
}

std::unique_ptr<AbstractCiphertext> CudaCiphertext::clone() const { return std::make_unique<CudaCiphertext>(*this); }

// ---- rotation (Evaluator::rotate_rows, SealCiphertext.cpp:52-61)
std::unique_ptr<AbstractCiphertext> CudaCiphertext::rotateRows(int steps) const {
  auto result = std::make_unique<CudaCiphertext>(getFactory());
  getFactory().checkResult(abc_rotate_rows(getFactory().context(), result->handle, handle, steps), result->handle);
  return result;
}
void CudaCiphertext::rotateRowsInplace(int steps) {
  getFactory().checkResult(abc_rotate_rows(getFactory().context(), handle, handle, steps), handle);
}

// ---- ctxt-ctxt (SealCiphertext.cpp:90-124)
std::unique_ptr<AbstractCiphertext> CudaCiphertext::add(const AbstractCiphertext &operand) const {
  auto result = std::make_unique<CudaCiphertext>(getFactory());
  getFactory().checkResult(abc_add(getFactory().context(), result->handle, handle, cast(operand).handle), result->handle);
  return result;
}
std::unique_ptr<AbstractCiphertext> CudaCiphertext::subtract(const AbstractCiphertext &operand) const {
  auto result = std::make_unique<CudaCiphertext>(getFactory());
  getFactory().checkResult(abc_sub(getFactory().context(), result->handle, handle, cast(operand).handle), result->handle);
  return result;
}
std::unique_ptr<AbstractCiphertext> CudaCiphertext::multiply(const AbstractCiphertext &operand) const {
  // multiply + relinearize_inplace in one call
  auto result = std::make_unique<CudaCiphertext>(getFactory());
  getFactory().checkResult(abc_mul_relin(getFactory().context(), result->handle, handle, cast(operand).handle), result->handle);
  return result;
}
void CudaCiphertext::addInplace(const AbstractCiphertext &operand) {
  getFactory().checkResult(abc_add(getFactory().context(), handle, handle, cast(operand).handle), handle);
}
void CudaCiphertext::subtractInplace(const AbstractCiphertext &operand) {
  getFactory().checkResult(abc_sub(getFactory().context(), handle, handle, cast(operand).handle), handle);
}
void CudaCiphertext::multiplyInplace(const AbstractCiphertext &operand) {
  getFactory().checkResult(abc_mul_relin(getFactory().context(), handle, handle, cast(operand).handle), handle);
}

// ---- ctxt-plain (SealCiphertext.cpp:130-202)
std::unique_ptr<AbstractCiphertext> CudaCiphertext::addPlain(const ICleartext &operand) const {
  auto data = widen(intCleartext(operand, "ADD").getData());
  auto result = std::make_unique<CudaCiphertext>(getFactory());
  getFactory().checkResult(abc_add_plain(getFactory().context(), result->handle, handle, data.data(), data.size(), 1), result->handle);
  return result;
}
std::unique_ptr<AbstractCiphertext> CudaCiphertext::subtractPlain(const ICleartext &operand) const {
  auto data = widen(intCleartext(operand, "SUB").getData());
  auto result = std::make_unique<CudaCiphertext>(getFactory());
  getFactory().checkResult(abc_sub_plain(getFactory().context(), result->handle, handle, data.data(), data.size(), 1), result->handle);
  return result;
}
std::unique_ptr<AbstractCiphertext> CudaCiphertext::multiplyPlain(const ICleartext &operand) const {
  const auto &cleartextInt = intCleartext(operand, "MULTIPLY");
  auto result = std::make_unique<CudaCiphertext>(getFactory());
  if (cleartextInt.allEqual(-1)) {  // negate fast path (SealCiphertext.cpp:156-157)
    getFactory().checkResult(abc_negate(getFactory().context(), result->handle, handle), result->handle);
  } else {
    auto data = widen(cleartextInt.getData());
    getFactory().checkResult(abc_mul_plain(getFactory().context(), result->handle, handle, data.data(), data.size(), 1), result->handle);
  }
  return result;
}
void CudaCiphertext::addPlainInplace(const ICleartext &operand) {
  auto data = widen(intCleartext(operand, "ADD").getData());
  getFactory().checkResult(abc_add_plain(getFactory().context(), handle, handle, data.data(), data.size(), 1), handle);
}
void CudaCiphertext::subtractPlainInplace(const ICleartext &operand) {
  auto data = widen(intCleartext(operand, "SUBTRACT").getData());
  getFactory().checkResult(abc_sub_plain(getFactory().context(), handle, handle, data.data(), data.size(), 1), handle);
}
void CudaCiphertext::multiplyPlainInplace(const ICleartext &operand) {
  const auto &cleartextInt = intCleartext(operand, "MULTIPLY");
  if (cleartextInt.allEqual(-1)) {
    getFactory().checkResult(abc_negate(getFactory().context(), handle, handle), handle);
  } else {
    auto data = widen(cleartextInt.getData());
    getFactory().checkResult(abc_mul_plain(getFactory().context(), handle, handle, data.data(), data.size(), 1), handle);
  }
}

// ---- AbstractValue dispatch (SealCiphertext.cpp:208-239)
void CudaCiphertext::add_inplace(const AbstractValue &other) {
  if (auto otherAsCiphertext = dynamic_cast<const CudaCiphertext *>(&other)) {
    addInplace(*otherAsCiphertext);
  } else if (auto otherAsCleartext = dynamic_cast<const ICleartext *>(&other)) {
    addPlainInplace(*otherAsCleartext);
  } else {
    throw std::runtime_error("Operation ADD only supported for (CudaCiphertext,CudaCiphertext) and (CudaCiphertext, ICleartext).");
  }
}
void CudaCiphertext::subtract_inplace(const AbstractValue &other) {
  if (auto otherAsCiphertext = dynamic_cast<const CudaCiphertext *>(&other)) {
    subtractInplace(*otherAsCiphertext);
  } else if (auto otherAsCleartext = dynamic_cast<const ICleartext *>(&other)) {
    subtractPlainInplace(*otherAsCleartext);
  } else {
    throw std::runtime_error("Operation SUBTRACT only supported for (CudaCiphertext,CudaCiphertext) and (CudaCiphertext, ICleartext).");
  }
}
void CudaCiphertext::multiply_inplace(const AbstractValue &other) {
  if (auto otherAsCiphertext = dynamic_cast<const CudaCiphertext *>(&other)) {
    multiplyInplace(*otherAsCiphertext);
  } else if (auto otherAsCleartext = dynamic_cast<const ICleartext *>(&other)) {
    multiplyPlainInplace(*otherAsCleartext);
  } else {
    throw std::runtime_error("Operation MULTIPLY only supported for (CudaCiphertext,CudaCiphertext) and (CudaCiphertext, ICleartext).");
  }
}

// ---- unsupported on ciphertexts (SealCiphertext.cpp:241-309)
#define ABC_UNSUPPORTED(NAME)                                                                              \
  void CudaCiphertext::NAME(const AbstractValue &) {                                                       \
    throw std::runtime_error("Operation " #NAME " not supported for (CudaCiphertext, ANY).");            \
  }
ABC_UNSUPPORTED(divide_inplace)
ABC_UNSUPPORTED(modulo_inplace)
ABC_UNSUPPORTED(logicalAnd_inplace)
ABC_UNSUPPORTED(logicalOr_inplace)
ABC_UNSUPPORTED(logicalLess_inplace)
ABC_UNSUPPORTED(logicalLessEqual_inplace)
ABC_UNSUPPORTED(logicalGreater_inplace)
ABC_UNSUPPORTED(logicalGreaterEqual_inplace)
ABC_UNSUPPORTED(logicalEqual_inplace)
ABC_UNSUPPORTED(logicalNotEqual_inplace)
ABC_UNSUPPORTED(bitwiseAnd_inplace)
ABC_UNSUPPORTED(bitwiseXor_inplace)
ABC_UNSUPPORTED(bitwiseOr_inplace)
#undef ABC_UNSUPPORTED

void CudaCiphertext::logicalNot_inplace() {
  throw std::runtime_error("Operation logicalNot_inplace not supported for (CudaCiphertext, ANY). "
                           "For an arithmetic negation, multiply_inplace by (-1) instead.");
}
void CudaCiphertext::bitwiseNot_inplace() {
  throw std::runtime_error("Operation bitwiseNot_inplace not supported for (CudaCiphertext, ANY). "
                           "For an arithmetic negation, multiply_inplace by (-1) instead.");
}
