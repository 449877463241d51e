// CudaCiphertext — B200 drop-in for ABC's SealCiphertext.
//
// Implements the 15 ciphertext virtuals + clone of AbstractCiphertext
// (/root/reference/include/ast_opt/runtime/AbstractCiphertext.h:27-98) and the 18 AbstractValue virtuals
// (include/ast_opt/runtime/AbstractValue.h:13-47), method for method like SealCiphertext
// (src/runtime/SealCiphertext.cpp).  The object owns one device handle (abc_ct*); every op enqueues
// kernels on the factory's stream and returns.
#ifndef ABC_B200_CPP_CUDACIPHERTEXT_H_
#define ABC_B200_CPP_CUDACIPHERTEXT_H_

#include <memory>

#include "ast_opt/runtime/AbstractCiphertext.h"
#include "CudaCiphertextFactory.h"

class CudaCiphertext : public AbstractCiphertext {
 private:
  abc_ct *handle = nullptr;

 public:
  explicit CudaCiphertext(const std::reference_wrapper<const CudaCiphertextFactory> cudaFactory);
  /// Takes ownership of an existing device handle.
  CudaCiphertext(const std::reference_wrapper<const CudaCiphertextFactory> cudaFactory, abc_ct *owned);
  ~CudaCiphertext() override;

  CudaCiphertext(const CudaCiphertext &other);                 // O(1): shares the device buffer, copy-on-write (abc_ct_clone)
  CudaCiphertext(CudaCiphertext &&other) noexcept;
  CudaCiphertext &operator=(const CudaCiphertext &other);
  CudaCiphertext &operator=(CudaCiphertext &&other);           // throws across factories (SealCiphertext.cpp:29-31)

  [[nodiscard]] abc_ct *getHandle() const { return handle; }

  /// Invariant noise budget in bits (SealCiphertext::noiseBits, src/runtime/SealCiphertext.cpp:80-83); the minimum
  /// over the instances of a batched factory.
  [[nodiscard]] int noiseBits() const;

  [[nodiscard]] std::unique_ptr<AbstractCiphertext> multiply(const AbstractCiphertext &operand) const override;
  void multiplyInplace(const AbstractCiphertext &operand) override;
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> add(const AbstractCiphertext &operand) const override;
  void addInplace(const AbstractCiphertext &operand) override;
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> subtract(const AbstractCiphertext &operand) const override;
  void subtractInplace(const AbstractCiphertext &operand) override;
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> rotateRows(int steps) const override;
  void rotateRowsInplace(int steps) override;
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> multiplyPlain(const ICleartext &operand) const override;
  void multiplyPlainInplace(const ICleartext &operand) override;
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> addPlain(const ICleartext &operand) const override;
  void addPlainInplace(const ICleartext &operand) override;
  [[nodiscard]] std::unique_ptr<AbstractCiphertext> subtractPlain(const ICleartext &operand) const override;
  void subtractPlainInplace(const ICleartext &operand) override;
  std::unique_ptr<AbstractCiphertext> clone() const override;

  void add_inplace(const AbstractValue &other) override;
  void subtract_inplace(const AbstractValue &other) override;
  void multiply_inplace(const AbstractValue &other) override;
  void divide_inplace(const AbstractValue &other) override;
  void modulo_inplace(const AbstractValue &other) override;
  void logicalAnd_inplace(const AbstractValue &other) override;
  void logicalOr_inplace(const AbstractValue &other) override;
  void logicalLess_inplace(const AbstractValue &other) override;
  void logicalLessEqual_inplace(const AbstractValue &other) override;
  void logicalGreater_inplace(const AbstractValue &other) override;
  void logicalGreaterEqual_inplace(const AbstractValue &other) override;
  void logicalEqual_inplace(const AbstractValue &other) override;
  void logicalNotEqual_inplace(const AbstractValue &other) override;
  void logicalNot_inplace() override;
  void bitwiseAnd_inplace(const AbstractValue &other) override;
  void bitwiseXor_inplace(const AbstractValue &other) override;
  void bitwiseOr_inplace(const AbstractValue &other) override;
  void bitwiseNot_inplace() override;

  [[nodiscard]] const CudaCiphertextFactory &getFactory() const override;
};

#endif  // ABC_B200_CPP_CUDACIPHERTEXT_H_
