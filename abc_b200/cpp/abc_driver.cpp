// abc_driver — runs ABC programs through the reference's own Parser / TypeCheckingVisitor / RuntimeVisitor
// (compiled unchanged from /root/reference/src) with CudaCiphertextFactory as the ciphertext backend.
//
//   abc_driver kats            the SEAL-backed cases of test/runtime/RuntimeVisitorTest.cpp and
//                              test/runtime/SealCiphertextFactoryTest.cpp with the factory swapped (N=4096)
//   abc_driver programs [N]    the hand-written batched programs of SURVEY.md 8(d) at N (default 8192), secret
//                              inputs: HammingDistance, L2Distance (n = N/2), BoxBlur, GxKernel (64x64 image),
//                              checked against plain C++ evaluations of the same functions
//   abc_driver programs N --batch B [--steps K]
//                              the lock-step batch driver: ONE RuntimeVisitor walk per step drives B
//                              independent instances of each program; the per-instance values of the `secret` inputs are
//                              registered with the factory (setBatchInputs), every instance is checked against the plain
//                              evaluation, and the L2Distance line reports the metric of bench.py's end-to-end leg
//                              (mul+relin + rotate ops/s, encrypt and decrypt included)
//   abc_driver demo <output_filename> [N] [--batch B]
//                              the reference's `ast_demo demo <output_filename>` (examples/main.cpp:33-46): one CSV row
//                              t_keygen,t_input_encryption,t_computation,t_decryption (ms) of the L2Distance ladder
// Prints one line per case and a JSON summary; exit code 1 on any mismatch.
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <functional>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "CudaCiphertext.h"
#include "CudaCiphertextFactory.h"
#include "ast_opt/parser/Parser.h"
#include "ast_opt/runtime/Cleartext.h"
#include "ast_opt/runtime/RuntimeVisitor.h"
#include "ast_opt/visitor/TypeCheckingVisitor.h"

namespace {
int failures = 0, cases = 0;

void report(const std::string &name, bool ok, const std::string &detail = "") {
  ++cases;
  if (!ok) ++failures;
  std::cout << (ok ? "[ ok ] " : "[FAIL] ") << name << (detail.empty() ? "" : "  " + detail) << std::endl;
}

std::string listOf(const std::vector<int> &v) {
  std::stringstream ss;
  ss << "{";
  for (size_t i = 0; i < v.size(); ++i) ss << (i ? "," : "") << v[i];
  ss << "}";
  return ss.str();
}

struct Var { std::string name; bool secret; };

using Results = std::unordered_map<std::string, std::vector<int64_t>>;

// The flow of test/runtime/RuntimeVisitorTest.cpp:224-262: parse inputs/program/outputs, pre-seed the type
// checker with the input variables, run, decrypt.
Results runProgram(CudaCiphertextFactory &factory, const std::string &inputs, const std::string &program,
                   const std::string &outputs, const std::vector<Var> &inputVars, double *seconds = nullptr) {
  auto astInput = Parser::parse(inputs);
  auto astProgram = Parser::parse(program);
  auto astOutput = Parser::parse(outputs);
  TypeCheckingVisitor tcv;
  auto rootScope = std::make_unique<Scope>(*astProgram);
  for (const auto &v : inputVars) {
    auto scopedIdentifier = std::make_unique<ScopedIdentifier>(*rootScope, v.name);
    rootScope->addIdentifier(v.name);
    tcv.addVariableDatatype(*scopedIdentifier, Datatype(Type::INT, v.secret));
  }
  tcv.setRootScope(std::move(rootScope));
  astProgram->accept(tcv);
  auto secretTaintedNodesMap = tcv.getSecretTaintedNodes();
  auto t0 = std::chrono::steady_clock::now();
  RuntimeVisitor srv(factory, *astInput, secretTaintedNodesMap);
  srv.executeAst(*astProgram);
  auto output = srv.getOutput(*astOutput);
  Results res;
  for (const auto &[identifier, value] : output) {
    std::vector<int64_t> plainValues;
    if (auto ciphertext = dynamic_cast<AbstractCiphertext *>(value.get())) {
      factory.decryptCiphertext(*ciphertext, plainValues);
    } else if (auto cleartextInt = dynamic_cast<Cleartext<int> *>(value.get())) {
      auto d = cleartextInt->getData();
      plainValues.assign(d.begin(), d.end());
    } else if (auto cleartextBool = dynamic_cast<Cleartext<bool> *>(value.get())) {
      auto d = cleartextBool->getData();
      plainValues.assign(d.begin(), d.end());
    } else {
      throw std::runtime_error("Could not determine type of result.");
    }
    res[identifier] = plainValues;
  }
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return res;
}

bool prefixEquals(const std::vector<int64_t> &got, const std::vector<int64_t> &want) {
  if (got.size() < want.size()) return false;
  for (size_t i = 0; i < want.size(); ++i)
    if (got[i] != want[i]) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------ kats
void checkCiphertextData(CudaCiphertextFactory &f, AbstractCiphertext &ct, const std::vector<int64_t> &expected,
                         const std::string &name) {
  // SealCiphertextFactoryTest.cpp:22-41: first values as given, every remaining slot = last value, size = N
  std::vector<int64_t> result;
  f.decryptCiphertext(ct, result);
  bool ok = result.size() == f.getCiphertextSlotSize() && prefixEquals(result, expected);
  for (size_t i = expected.size(); ok && i < result.size(); ++i) ok = result[i] == expected.back();
  report(name, ok);
}

void runKats() {
  CudaCiphertextFactory f(4096);
  const std::vector<int64_t> data1 = {3, 3, 1, 4, 5, 9}, data2 = {0, 1, 2, 1, 10, 21};
  const std::vector<int64_t> sum = {3, 4, 3, 5, 15, 30}, diff = {3, 2, -1, 3, -5, -12}, prod = {0, 3, 2, 4, 50, 189};
  {
    auto c = f.createCiphertext(data1);
    checkCiphertextData(f, *c, data1, "factory.createCiphertext");
    auto c1 = f.createCiphertext(data1), c2 = f.createCiphertext(data2);
    auto r = c1->add(*c2); checkCiphertextData(f, *r, sum, "factory.add");
    r = c1->subtract(*c2); checkCiphertextData(f, *r, diff, "factory.sub");
    r = c1->multiply(*c2); checkCiphertextData(f, *r, prod, "factory.multiply");
    checkCiphertextData(f, *c1, data1, "factory.operand1 unchanged");
    checkCiphertextData(f, *c2, data2, "factory.operand2 unchanged");
    auto a = c1->clone(); a->addInplace(*c2); checkCiphertextData(f, *a, sum, "factory.addInplace");
    a = c1->clone(); a->subtractInplace(*c2); checkCiphertextData(f, *a, diff, "factory.subInplace");
    a = c1->clone(); a->multiplyInplace(*c2); checkCiphertextData(f, *a, prod, "factory.multiplyInplace");
    Cleartext<int> pt(std::vector<int>{0, 1, 2, 1, 10, 21});
    r = c1->addPlain(pt); checkCiphertextData(f, *r, sum, "factory.addPlain");
    r = c1->subtractPlain(pt); checkCiphertextData(f, *r, diff, "factory.subPlain");
    r = c1->multiplyPlain(pt); checkCiphertextData(f, *r, prod, "factory.multiplyPlain");
    a = c1->clone(); a->addPlainInplace(pt); checkCiphertextData(f, *a, sum, "factory.addPlainInplace");
    a = c1->clone(); a->subtractPlainInplace(pt); checkCiphertextData(f, *a, diff, "factory.subPlainInplace");
    a = c1->clone(); a->multiplyPlainInplace(pt); checkCiphertextData(f, *a, prod, "factory.multiplyPlainInplace");
    Cleartext<int> minusOne(std::vector<int>{-1});
    r = c1->multiplyPlain(minusOne); checkCiphertextData(f, *r, {-3, -3, -1, -4, -5, -9}, "factory.multiplyPlain(-1) negates");
    // rotation incl. wrap-around at N/2 (SealCiphertextFactoryTest.cpp:51-140)
    const std::vector<int64_t> rd = {123456, 3, 1, 4, 5, 9, 5, 2, 1, 5};
    auto rc = f.createCiphertext(rd);
    const size_t half = f.getCiphertextSlotSize() / 2;
    for (int steps : {4, -24}) {
      auto rot = rc->rotateRows(steps);
      std::vector<int64_t> dv, full(rd);
      full.resize(f.getCiphertextSlotSize(), rd.back());
      f.decryptCiphertext(*rot, dv);
      bool ok = true;
      for (size_t row = 0; row < 2; ++row)
        for (size_t i = 0; i < half; ++i)
          ok = ok && dv[row * half + i] == full[row * half + ((i + half + steps) % half)];
      report("factory.rotateRows(" + std::to_string(steps) + ")", ok);
    }
    bool threw = false;
    try { auto big = f.createCiphertext(std::vector<int64_t>(4097, 1)); } catch (std::runtime_error &) { threw = true; }
    report("factory.createCiphertext(> N values) throws std::runtime_error", threw);
  }
  const std::vector<Var> in0 = {{"__input0__", true}}, in01 = {{"__input0__", true}, {"__input1__", true}};
  {
    auto r = runProgram(f, "secret int __input0__ = {43, 1, 1, 1, 22, 11, 425, 0, 1, 7};",
                        "__input0__ = rotate(__input0__, -4);", "y = __input0__;", in0);
    report("visitor.testRotateNegative", prefixEquals(r["y"], {7, 7, 7, 7, 43, 1, 1, 1, 22, 11, 425, 0, 1, 7}));
    r = runProgram(f, "secret int __input0__ = {43, 1, 1, 1, 22, 11, 425, 0, 1, 7};",
                   "__input0__ = rotate(__input0__, 6);", "y = __input0__;", in0);
    report("visitor.testRotatePositive", prefixEquals(r["y"], {425, 0, 1, 7, 7, 7, 7, 7, 7}));
    r = runProgram(f,
                   "secret int __input0__ = {43,  1,   1,   1,  22, 11, 425,  0, 1, 7};\n"
                   "secret int __input1__ = {24, 34, 222,   4,    1, 4,   9, 22, 1, 3};",
                   "secret int result = __input0__ *** __input1__;\nreturn result;", "y = result;", in01);
    report("visitor.testFheMultCtxtCtxt", prefixEquals(r["y"], {1032, 34, 222, 4, 22, 44, 3825, 0, 1, 21}));
    r = runProgram(f, "secret int __input0__ = {43,  1,   1,  22, 11, 7};",
                   "int i = 19;\nsecret int result = __input0__ *** i;\nreturn result;", "y = result;\nx = result[3];", in0);
    report("visitor.testFheMultCtxtPlain", prefixEquals(r["y"], {817, 19, 19, 418, 209, 133}) && prefixEquals(r["x"], {418}));
    r = runProgram(f, "secret int __input0__ = {43,  1,   1,  22, 11, 7};",
                   "int i = 19;\nsecret int result = i *** __input0__;\nreturn result;", "y = result;\nx = result[3];", in0);
    report("visitor.testFheMultPlainCtxt", prefixEquals(r["y"], {817, 19, 19, 418, 209, 133}) && prefixEquals(r["x"], {418}));
    r = runProgram(f, "secret int __input0__ = {43, 1, 1, 1, 22, 11, 425, 0, 1, 7};",
                   "int LIMIT = 10;\nsecret int result = 0;\nfor (int i = 0; i < LIMIT; i = i + 1) {\n"
                   "  result = result + __input0__;\n}\nreturn;",
                   "y = result;", in0);
    report("visitor.testForLoop", prefixEquals(r["y"], {430, 10, 10, 10, 220, 110, 4250, 0, 10, 70}));
    bool threw = false;
    try {
      runProgram(f, "secret int sum = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10};", "secret int result = sum / sum;\nreturn result;",
                 "y = result;", {{"sum", true}});
    } catch (std::runtime_error &) { threw = true; }
    report("visitor.unsupported '/' on ciphertexts throws std::runtime_error", threw);
  }
  {
    // SealCiphertextFactory::createPlaintext x 3 (SealCiphertextFactory.h:95-107) + encryption of the result
    auto p1 = f.createPlaintext(std::vector<int64_t>{3, 3, 1, 4, 5, 9});
    auto p2 = f.createPlaintext(std::vector<int>{0, 1, 2, 1, 10, 21});
    auto p3 = f.createPlaintext((int64_t)-7);
    auto c1 = f.encryptPlaintext(*p1), c2 = f.encryptPlaintext(*p2), c3 = f.encryptPlaintext(*p3);
    checkCiphertextData(f, *c1, data1, "factory.createPlaintext(vector<int64_t>) + encrypt");
    checkCiphertextData(f, *c2, data2, "factory.createPlaintext(vector<int>) + encrypt");
    checkCiphertextData(f, *c3, {-7}, "factory.createPlaintext(int64_t) + encrypt");
  }
  {
    // all 15 operations a ciphertext rejects (SealCiphertext.cpp:241-309) throw the same exception type
    auto c1 = f.createCiphertext(data1), c2 = f.createCiphertext(data2);
    using BinOp = void (AbstractValue::*)(const AbstractValue &);
    const std::vector<std::pair<const char *, BinOp>> binary = {
        {"divide", &AbstractValue::divide_inplace}, {"modulo", &AbstractValue::modulo_inplace},
        {"logicalAnd", &AbstractValue::logicalAnd_inplace}, {"logicalOr", &AbstractValue::logicalOr_inplace},
        {"logicalLess", &AbstractValue::logicalLess_inplace}, {"logicalLessEqual", &AbstractValue::logicalLessEqual_inplace},
        {"logicalGreater", &AbstractValue::logicalGreater_inplace},
        {"logicalGreaterEqual", &AbstractValue::logicalGreaterEqual_inplace},
        {"logicalEqual", &AbstractValue::logicalEqual_inplace}, {"logicalNotEqual", &AbstractValue::logicalNotEqual_inplace},
        {"bitwiseAnd", &AbstractValue::bitwiseAnd_inplace}, {"bitwiseXor", &AbstractValue::bitwiseXor_inplace},
        {"bitwiseOr", &AbstractValue::bitwiseOr_inplace}};
    Cleartext<int> pt(std::vector<int>{1, 2, 3});
    for (const auto &[name, op] : binary) {
      int threw = 0;
      try { ((*c1).*op)(*c2); } catch (std::runtime_error &) { ++threw; }
      try { ((*c1).*op)(pt); } catch (std::runtime_error &) { ++threw; }
      report(std::string("unsupported.") + name + "_inplace(ciphertext | cleartext) throws std::runtime_error", threw == 2);
    }
    int threw = 0;
    try { c1->logicalNot_inplace(); } catch (std::runtime_error &) { ++threw; }
    try { c1->bitwiseNot_inplace(); } catch (std::runtime_error &) { ++threw; }
    report("unsupported.logicalNot_inplace / bitwiseNot_inplace throw std::runtime_error", threw == 2);
    checkCiphertextData(f, *c1, data1, "unsupported.operand unchanged after 15 rejected operations");
  }
  {
    // Cleartext<int>::subtract_inplace(ciphertext) (include/ast_opt/runtime/Cleartext.h:349-360) encrypts
    // the cleartext through the ciphertext's factory, subtracts, and DISCARDS the result: the cleartext is unchanged.  The
    // quirk is the reference's; what is checked is that it runs through this factory (a ciphertext is created and a
    // subtraction is launched) and leaves both operands as they were.
    Cleartext<int> plain(std::vector<int>{50, 60, 70, 80, 90, 100});
    auto c = f.createCiphertext(data1);
    const uint64_t before = f.launchCount();
    plain.subtract_inplace(*c);
    const uint64_t used = f.launchCount() - before;
    const std::vector<int> want = {50, 60, 70, 80, 90, 100};
    report("a16.Cleartext<int>::subtract_inplace(ciphertext): cleartext unchanged, encrypt + sub launched through the factory",
           plain.getData() == want && used >= 2, "launches " + std::to_string(used));
    checkCiphertextData(f, *c, data1, "a16.ciphertext operand unchanged");
    // the same through the interpreter: `i --- x` with a plain left operand keeps the cleartext (RuntimeVisitor.cpp:69-71),
    // which the secret declaration then encrypts
    auto r = runProgram(f, "secret int __input0__ = {43,  1,   1,  22, 11, 7};",
                        "int i = 19;\nsecret int result = i --- __input0__;\nreturn result;", "y = result;", in0);
    report("a16.visitor `plain --- cipher` yields the (encrypted) cleartext operand", prefixEquals(r["y"], {19, 19, 19, 19, 19, 19}));
    // and the well-defined direction for comparison
    r = runProgram(f, "secret int __input0__ = {43,  1,   1,  22, 11, 7};",
                   "int i = 19;\nsecret int result = __input0__ --- i;\nreturn result;", "y = result;", in0);
    report("a16.visitor `cipher --- plain`", prefixEquals(r["y"], {24, -18, -18, 3, -8, -12}));
  }
  {
    // two factories built from the same caller-supplied 32-byte sampler key hold the same secret key: a ciphertext saved by
    // one (SEAL stream) decrypts under the other; a factory with an OS-drawn key cannot decrypt it
    std::array<uint8_t, 32> rngKey{};
    for (size_t i = 0; i < rngKey.size(); ++i) rngKey[i] = (uint8_t)(101 * i + 7);
    CudaCiphertextFactory a(4096, 0, 1, rngKey), b(4096, 0, 1, rngKey), other(4096);
    auto ct = a.createCiphertext(data1);
    auto stream = a.saveCiphertext(*ct);
    auto viaB = b.loadCiphertext(stream);
    checkCiphertextData(b, *viaB, data1, "rngkey.same 32-byte sampler key on two factories: ciphertext of one decrypts under the other");
    auto viaOther = other.loadCiphertext(stream);
    std::vector<int64_t> garbage;
    other.decryptCiphertext(*viaOther, garbage);
    report("rngkey.a factory with an OS-drawn key does not decrypt it", !prefixEquals(garbage, data1));
  }
  {
    // SEAL (SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT, its default) throws std::logic_error on a result whose c1 is
    // zero.  Off by default here (a host synchronisation per op); setThrowOnTransparent(true) mirrors it.
    auto c = f.createCiphertext(data1);
    auto d = c->clone();
    auto z = c->subtract(*d);
    checkCiphertextData(f, *z, {0, 0, 0, 0, 0, 0}, "transparent.default: x --- x is an encryption of zero, no throw");
    f.setThrowOnTransparent(true);
    int threw = 0;
    try { c->subtract(*d); } catch (std::logic_error &e) { threw += std::string(e.what()) == "result ciphertext is transparent"; }
    try { auto a = c->clone(); a->subtractInplace(*d); } catch (std::logic_error &) { ++threw; }
    try { Cleartext<int> zero(std::vector<int>{0}); c->multiplyPlain(zero); } catch (std::logic_error &) { ++threw; }
    report("transparent.mirrored: x --- x, subtractInplace, x *** 0 throw std::logic_error(\"result ciphertext is transparent\")", threw == 3);
    bool ok = true;
    try {
      auto r = c->add(*d); auto m = c->multiply(*d); auto t = c->rotateRows(3);
      Cleartext<int> two(std::vector<int>{2}); auto p = c->multiplyPlain(two);
    } catch (std::logic_error &) { ok = false; }
    report("transparent.mirrored: ordinary results pass the check", ok);
    // through the interpreter: RuntimeVisitor clones every variable read (RuntimeVisitor.cpp:436), so `x --- x` reaches it
    threw = 0;
    try {
      runProgram(f, "secret int __input0__ = {43,  1,   1,  22, 11, 7};", "secret int result = __input0__ --- __input0__;\nreturn result;",
                 "y = result;", in0);
    } catch (std::logic_error &) { ++threw; }
    report("transparent.mirrored: visitor `x --- x` throws std::logic_error", threw == 1);
    f.setThrowOnTransparent(false);
  }
}

// ------------------------------------------------------------------------------------------------ lock-step batch
// One interpreter walk for `batch` instances: the ASTs are parsed and type-checked once (outside the clock); a step =
// RuntimeVisitor construction (evaluates the input declarations: createCiphertext pulls every instance's values from the
// registered tables), executeAst, getOutput, decryptCiphertextBatchPinned.
struct BatchProgram {
  std::unique_ptr<AbstractNode> astInput, astProgram, astOutput;
  SecretTaintedNodesMap tainted;
  BatchProgram(const std::string &inputs, const std::string &program, const std::string &outputs, const std::vector<Var> &inputVars) {
    astInput = Parser::parse(inputs);
    astProgram = Parser::parse(program);
    astOutput = Parser::parse(outputs);
    TypeCheckingVisitor tcv;
    auto rootScope = std::make_unique<Scope>(*astProgram);
    for (const auto &v : inputVars) {
      auto scopedIdentifier = std::make_unique<ScopedIdentifier>(*rootScope, v.name);
      rootScope->addIdentifier(v.name);
      tcv.addVariableDatatype(*scopedIdentifier, Datatype(Type::INT, v.secret));
    }
    tcv.setRootScope(std::move(rootScope));
    astProgram->accept(tcv);
    tainted = tcv.getSecretTaintedNodes();
  }
  // returns the factory's pinned staging: batch * N decrypted slots of the (single) ciphertext output
  const int64_t *step(CudaCiphertextFactory &factory, bool async = false) {
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    factory.rewindBatchInputs();
    RuntimeVisitor srv(factory, *astInput, tainted);
    const auto t1 = now();
    srv.executeAst(*astProgram);
    const auto t2 = now();
    auto output = srv.getOutput(*astOutput);
    const auto t3 = now();
    for (const auto &[identifier, value] : output)
      if (auto ciphertext = dynamic_cast<AbstractCiphertext *>(value.get())) {
        const int64_t *r = async ? factory.decryptCiphertextBatchPinnedAsync(*ciphertext) : factory.decryptCiphertextBatchPinned(*ciphertext);
        if (getenv("ABC_DRIVER_TIMING"))
          std::cout << "  host ms: inputs+encrypt (enqueue) " << ms(t0, t1) << ", executeAst (enqueue) " << ms(t1, t2)
                    << ", getOutput " << ms(t2, t3) << ", decrypt (sync) " << ms(t3, now()) << std::endl;
        return r;
      }
    throw std::runtime_error("batch program: no ciphertext output");
  }
};

// ------------------------------------------------------------------------------------------------ programs
std::vector<int> randomVector(size_t n, unsigned seed) {
  // the reference's generator: std::default_random_engine(4673838), uniform ints in [0,1024]
  // (test/end-to-end/BoxBlurTest.cpp:111-124)
  std::default_random_engine engine(seed);
  std::uniform_int_distribution<int> dist(0, 1024);
  std::vector<int> v(n);
  for (auto &x : v) x = dist(engine);
  return v;
}

// plain evaluation of a 3x3 wrap-around stencil, pixel (x,y) at x*size + y
std::vector<int64_t> stencil(const std::vector<int> &img, int size, const int w[3][3]) {
  std::vector<int64_t> out(img.size());
  const long total = (long)img.size();
  for (int x = 0; x < size; ++x)
    for (int y = 0; y < size; ++y) {
      int64_t acc = 0;
      for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
          long idx = (((long)(x + dx) * size + (y + dy)) % total + total) % total;
          acc += (int64_t)w[dx + 1][dy + 1] * img[idx];
        }
      out[(size_t)x * size + y] = acc;
    }
  return out;
}

// batched form: one rotation per non-zero tap, ciphertext on the left of every binary op
std::string stencilProgram(int size, const int w[3][3]) {
  std::stringstream p;
  // acc starts as the first tap (not as img --- img: SEAL rejects that transparent result)
  p << "secret int acc = img;\nsecret int r = img;\n";
  bool first = true;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy) {
      const int weight = w[dx + 1][dy + 1], k = dx * size + dy;
      if (weight == 0) continue;
      if (k == 0) p << "r = img;\n";
      else p << "r = rotate(img, " << k << ");\n";
      if (std::abs(weight) != 1) p << "r = r *** " << std::abs(weight) << ";\n";
      if (first && weight > 0) p << "acc = r;\n";
      else p << "acc = acc " << (weight > 0 ? "+++" : "---") << " r;\n";
      first = false;
    }
  p << "return acc;\n";
  return p.str();
}

void runPrograms(unsigned N) {
  CudaCiphertextFactory f(N);
  const size_t n = N / 2;  // one batching row
  const int64_t t = N <= 8192 ? 1032193 : 786433;
  auto centre = [&](int64_t v) { v %= t; if (v < 0) v += t; return v > t / 2 ? v - t : v; };
  auto x = randomVector(n, 4673838), y = randomVector(n, 4673839);
  std::vector<int> bx(n), by(n);
  for (size_t i = 0; i < n; ++i) { bx[i] = x[i] & 1; by[i] = y[i] & 1; }

  // rotate-and-sum ladder: log2(n) rotations, result in slot 0 (and every slot of the row)
  std::stringstream ladder;
  ladder << "secret int d = x --- y;\nsecret int s = d *** d;\n";
  for (size_t k = n / 2; k >= 1; k /= 2) ladder << "s = s +++ rotate(s, " << k << ");\n";
  ladder << "return s;\n";
  const std::vector<Var> xy = {{"x", true}, {"y", true}};
  double secs = 0;
  {
    auto r = runProgram(f, "secret int x = " + listOf(bx) + ";\nsecret int y = " + listOf(by) + ";", ladder.str(), "s = s;", xy, &secs);
    int64_t want = 0;
    for (size_t i = 0; i < n; ++i) want += (bx[i] - by[i]) * (bx[i] - by[i]);
    report("program.HammingDistance n=" + std::to_string(n), !r["s"].empty() && r["s"][0] == centre(want),
           "expected " + std::to_string(centre(want)) + " got " + std::to_string(r["s"].empty() ? -1 : r["s"][0]) +
               " in " + std::to_string(secs) + " s");
  }
  {
    auto r = runProgram(f, "secret int x = " + listOf(x) + ";\nsecret int y = " + listOf(y) + ";", ladder.str(), "s = s;", xy, &secs);
    int64_t want = 0;
    for (size_t i = 0; i < n; ++i) want += (int64_t)(x[i] - y[i]) * (x[i] - y[i]);
    report("program.L2Distance n=" + std::to_string(n), !r["s"].empty() && r["s"][0] == centre(want),
           "expected " + std::to_string(centre(want)) + " got " + std::to_string(r["s"].empty() ? -1 : r["s"][0]) +
               " in " + std::to_string(secs) + " s");
  }
  // the reference's wrap-around index is exactly a cyclic row rotation only when the image fills the row
  if (const int size = (int)std::lround(std::sqrt((double)n)); (size_t)size * size == n) {
    auto img = randomVector((size_t)size * size, 4673838);
    const int box[3][3] = {{1, 1, 1}, {1, 1, 1}, {1, 1, 1}};
    const int gx[3][3] = {{1, 2, 1}, {0, 0, 0}, {-1, -2, -1}};  // weightMatrix of GxKernelTest.cpp:22
    for (int which = 0; which < 2; ++which) {
      const auto &w = which == 0 ? box : gx;
      auto r = runProgram(f, "secret int img = " + listOf(img) + ";", stencilProgram(size, w), "acc = acc;", {{"img", true}}, &secs);
      auto want = stencil(img, size, w);
      bool ok = r["acc"].size() >= want.size();
      for (size_t i = 0; ok && i < want.size(); ++i) ok = r["acc"][i] == centre(want[i]);
      report(std::string("program.") + (which == 0 ? "BoxBlur " : "GxKernel ") + std::to_string(size) + "x" + std::to_string(size), ok,
             "in " + std::to_string(secs) + " s");
    }
  }
  std::cout << "launches=" << f.launchCount() << std::endl;
}
void runProgramsBatch(unsigned N, unsigned B, int steps) {
  CudaCiphertextFactory f(N, 0, B, 0);
  const size_t n = N / 2;
  const int64_t t = N <= 8192 ? 1032193 : 786433;
  auto centre = [&](int64_t v) { v %= t; if (v < 0) v += t; return v > t / 2 ? v - t : v; };
  // per-instance inputs, instance-major tables
  std::vector<std::vector<int>> xs(B), ys(B);
  std::vector<int64_t> tx((size_t)B * n), ty((size_t)B * n);
  for (unsigned b = 0; b < B; ++b) {
    xs[b] = randomVector(n, 4673838 + 2 * b); ys[b] = randomVector(n, 4673839 + 2 * b);
    for (size_t i = 0; i < n; ++i) { tx[b * n + i] = xs[b][i]; ty[b * n + i] = ys[b][i]; }
  }
  std::stringstream ladder;
  ladder << "secret int d = x --- y;\nsecret int s = d *** d;\n";
  int rotations = 0;
  for (size_t k = n / 2; k >= 1; k /= 2) { ladder << "s = s +++ rotate(s, " << k << ");\n"; ++rotations; }
  ladder << "return s;\n";
  const std::vector<Var> xy = {{"x", true}, {"y", true}};
  {
    // the literals are placeholders: every instance's values come from the registered tables
    BatchProgram prog("secret int x = {0};\nsecret int y = {0};", ladder.str(), "s = s;", xy);
    f.setBatchInputs({tx, ty});                  // registered (and page-locked) once; every walk rewinds to the first table
    const int64_t *out = prog.step(f);           // warm-up (scratch growth, key copies) + check of every instance
    bool ok = true;
    for (unsigned b = 0; b < B && ok; ++b) {
      int64_t want = 0;
      for (size_t i = 0; i < n; ++i) want += (int64_t)(xs[b][i] - ys[b][i]) * (xs[b][i] - ys[b][i]);
      ok = out[(size_t)b * N] == centre(want);
    }
    prog.step(f);                                // second warm-up walk: the buffer free list reaches its steady size
    f.synchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int s_ = 0; s_ < steps; ++s_) out = prog.step(f, true);   // the D2H of walk i runs under walk i + 1
    f.waitDecryptions();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (unsigned b = 0; b < B && ok; ++b) {                          // the last walk's result, every instance again
      int64_t want = 0;
      for (size_t i = 0; i < n; ++i) want += (int64_t)(xs[b][i] - ys[b][i]) * (xs[b][i] - ys[b][i]);
      ok = out[(size_t)b * N] == centre(want);
    }
    const double ops = (double)B * (1 + rotations) * steps / secs;
    report("batch.L2Distance n=" + std::to_string(n) + " B=" + std::to_string(B), ok,
           std::to_string(secs / steps * 1e3) + " ms per walk, " + std::to_string((long)ops) + " mul+relin & rotate ops/s end to end");
    std::cout << "{\"program\": \"L2Distance\", \"N\": " << N << ", \"batch\": " << B << ", \"steps\": " << steps
              << ", \"ms_per_walk\": " << secs / steps * 1e3 << ", \"ops_per_s\": " << ops
              << ", \"instances_per_s\": " << (double)B * steps / secs << ", \"all_instances_checked\": " << (ok ? "true" : "false") << "}"
              << std::endl;
  }
  if (const int size = (int)std::lround(std::sqrt((double)n)); (size_t)size * size == n) {
    const int box[3][3] = {{1, 1, 1}, {1, 1, 1}, {1, 1, 1}};
    const int gx[3][3] = {{1, 2, 1}, {0, 0, 0}, {-1, -2, -1}};
    std::vector<std::vector<int>> imgs(B);
    std::vector<int64_t> ti((size_t)B * n);
    for (unsigned b = 0; b < B; ++b) {
      imgs[b] = randomVector(n, 99 + b);
      for (size_t i = 0; i < n; ++i) ti[b * n + i] = imgs[b][i];
    }
    for (int which = 0; which < 2; ++which) {
      const auto &w = which == 0 ? box : gx;
      BatchProgram prog("secret int img = {0};", stencilProgram(size, w), "acc = acc;", {{"img", true}});
      f.setBatchInputs({ti});
      const int64_t *out = prog.step(f);
      bool ok = true;
      for (unsigned b = 0; b < B && ok; b += std::max(1u, B / 16)) {   // every 1/16th instance in full
        auto want = stencil(imgs[b], size, w);
        for (size_t i = 0; ok && i < want.size(); ++i) ok = out[(size_t)b * N + i] == centre(want[i]);
      }
      prog.step(f);                              // second warm-up walk: the buffer free list reaches its steady size
      f.synchronize();
      auto t0 = std::chrono::steady_clock::now();
      for (int s_ = 0; s_ < steps; ++s_) out = prog.step(f);
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      report(std::string("batch.") + (which == 0 ? "BoxBlur " : "GxKernel ") + std::to_string(size) + "x" + std::to_string(size) +
                 " B=" + std::to_string(B), ok,
             std::to_string(secs / steps * 1e3) + " ms per walk, " + std::to_string((long)(B * steps / secs)) + " instances/s end to end");
    }
  }
  std::cout << "launches=" << f.launchCount() << std::endl;
}

// ------------------------------------------------------------------------------------------------ demo
// The reference's `ast_demo demo <output_filename>` (examples/main.cpp:33-46) writes one CSV row
// `t_keygen,t_input_encryption,t_computation,t_decryption` (ms; a placeholder row there).  Here the row is measured: the
// L2Distance ladder through the unmodified RuntimeVisitor on the CUDA factory, the stream drained after every phase.
void runDemo(const std::string &filename, unsigned N, unsigned B) {
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const size_t n = N / 2;
  const int64_t t = N <= 8192 ? 1032193 : 786433;
  auto centre = [&](int64_t v) { v %= t; if (v < 0) v += t; return v > t / 2 ? v - t : v; };
  std::vector<int64_t> tx((size_t)B * n), ty((size_t)B * n);
  for (unsigned b = 0; b < B; ++b) {
    auto x = randomVector(n, 4673838 + 2 * b), y = randomVector(n, 4673839 + 2 * b);
    for (size_t i = 0; i < n; ++i) { tx[b * n + i] = x[i]; ty[b * n + i] = y[i]; }
  }
  std::stringstream ladder;
  ladder << "secret int d = x --- y;\nsecret int s = d *** d;\n";
  for (size_t k = n / 2; k >= 1; k /= 2) ladder << "s = s +++ rotate(s, " << k << ");\n";
  ladder << "return s;\n";
  BatchProgram prog("secret int x = {0};\nsecret int y = {0};", ladder.str(), "s = s;", {{"x", true}, {"y", true}});
  const auto t0 = now();
  CudaCiphertextFactory f(N, 0, B, 0);          // setupSealContext: parameters, sk / pk / relin / Galois keys
  f.synchronize();
  const auto t1 = now();
  f.setBatchInputs({tx, ty});
  double phase[3] = {0, 0, 0};
  bool ok = true;
  for (int rep = 0; rep < 2; ++rep) {            // the second walk is the one reported (the first grows the scratch slots)
    f.rewindBatchInputs();
    const auto a = now();
    RuntimeVisitor srv(f, *prog.astInput, prog.tainted);
    f.synchronize();
    const auto b = now();
    srv.executeAst(*prog.astProgram);
    f.synchronize();
    const auto c = now();
    auto output = srv.getOutput(*prog.astOutput);
    const int64_t *out = nullptr;
    for (const auto &[identifier, value] : output)
      if (auto ciphertext = dynamic_cast<AbstractCiphertext *>(value.get())) out = f.decryptCiphertextBatchPinned(*ciphertext);
    const auto d = now();
    phase[0] = ms(a, b); phase[1] = ms(b, c); phase[2] = ms(c, d);
    for (unsigned i = 0; i < B && ok && out; ++i) {
      int64_t want = 0;
      for (size_t j = 0; j < n; ++j) want += (tx[i * n + j] - ty[i * n + j]) * (tx[i * n + j] - ty[i * n + j]);
      ok = out[(size_t)i * N] == centre(want);
    }
    ok = ok && out;
  }
  std::ofstream file(filename);
  file << "t_keygen,t_input_encryption,t_computation,t_decryption\n"
       << ms(t0, t1) << "," << phase[0] << "," << phase[1] << "," << phase[2] << std::endl;
  report("demo.L2Distance N=" + std::to_string(N) + " B=" + std::to_string(B) + " -> " + filename, ok && file.good(),
         "ms: keygen " + std::to_string(ms(t0, t1)) + ", input encryption " + std::to_string(phase[0]) + ", computation " +
             std::to_string(phase[1]) + ", decryption " + std::to_string(phase[2]));
}
}  // namespace

int main(int argc, char **argv) {
  const std::string mode = argc > 1 ? argv[1] : "kats";
  try {
    if (mode == "kats") runKats();
    else if (mode == "programs") {
      const unsigned N = argc > 2 ? (unsigned)std::stoul(argv[2]) : 8192;
      unsigned batch = 0; int steps = 5;
      for (int i = 3; i + 1 < argc; i += 2) {
        if (std::string(argv[i]) == "--batch") batch = (unsigned)std::stoul(argv[i + 1]);
        if (std::string(argv[i]) == "--steps") steps = std::stoi(argv[i + 1]);
      }
      if (batch) runProgramsBatch(N, batch, steps);
      else runPrograms(N);
    }
    else if (mode == "demo") {
      if (argc < 3) { std::cerr << "usage: abc_driver demo <output_filename> [N] [--batch B]" << std::endl; return 2; }
      const unsigned N = argc > 3 ? (unsigned)std::stoul(argv[3]) : 8192;
      unsigned batch = 1;
      for (int i = 4; i + 1 < argc; i += 2)
        if (std::string(argv[i]) == "--batch") batch = (unsigned)std::stoul(argv[i + 1]);
      runDemo(argv[2], N, batch);
    }
    else { std::cerr << "usage: abc_driver kats | programs [N] [--batch B --steps K] | demo <output_filename> [N] [--batch B]" << std::endl; return 2; }
  } catch (const std::exception &e) {
    std::cout << "[FAIL] uncaught exception: " << e.what() << std::endl;
    ++failures;
  }
  std::cout << "{\"mode\": \"" << mode << "\", \"cases\": " << cases << ", \"failures\": " << failures << "}" << std::endl;
  return failures ? 1 : 0;
}
