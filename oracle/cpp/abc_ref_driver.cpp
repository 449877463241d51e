// abc_ref_driver — the CPU baseline arm of bench.py (TEST INFRASTRUCTURE): the batched L2Distance program of
// SURVEY.md 8(d) through the reference's own Parser / TypeCheckingVisitor / RuntimeVisitor (compiled unchanged from
// /root/reference/src into oracle/_ref/libabc_ref.a) with OracleCiphertextFactory as the ciphertext backend, one
// interpreter walk per instance, T host workers.  The workers are forked PROCESSES, not threads: the reference's parser and
// AST keep global state (unique node ids, ...) that is not thread-safe, while the oracle context (parameters, keys) is
// read-only after key generation and is shared copy-on-write.
//   abc_ref_driver N threads instances_per_thread steps warmup
// Prints one JSON line: mul+relin & rotate ops/s (encrypt x, y + program + decrypt per instance), result check.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <random>
#include <sstream>
#include <thread>
#include <vector>

#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include "OracleCiphertextFactory.hpp"
#include "ast_opt/parser/Parser.h"
#include "ast_opt/runtime/RuntimeVisitor.h"
#include "ast_opt/visitor/TypeCheckingVisitor.h"

int main(int argc, char **argv) {
  const unsigned N = argc > 1 ? std::stoul(argv[1]) : 8192;
  const int T = argc > 2 ? std::stoi(argv[2]) : (int)std::thread::hardware_concurrency();
  const int per = argc > 3 ? std::stoi(argv[3]) : 2, steps = argc > 4 ? std::stoi(argv[4]) : 1, warmup = argc > 5 ? std::stoi(argv[5]) : 0;
  const size_t n = N / 2;
  const int64_t t = N <= 8192 ? 1032193 : 786433;
  auto centre = [&](int64_t v) { v %= t; if (v < 0) v += t; return v > t / 2 ? v - t : v; };
  OracleCiphertextFactory base(N);
  std::stringstream ladder;
  ladder << "secret int d = x --- y;\nsecret int s = d *** d;\n";
  int rotations = 0;
  for (size_t k = n / 2; k >= 1; k /= 2) { ladder << "s = s +++ rotate(s, " << k << ");\n"; ++rotations; }
  ladder << "return s;\n";
  const int total = T * per;
  std::vector<std::vector<int64_t>> xs(total), ys(total);
  std::vector<int64_t> want(total);
  // results come back through anonymous shared memory
  int64_t *got = static_cast<int64_t *>(mmap(nullptr, sizeof(int64_t) * total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
  for (int i = 0; i < total; ++i) got[i] = 0;
  for (int i = 0; i < total; ++i) {
    std::default_random_engine engine(4673838 + i);
    std::uniform_int_distribution<int> dist(0, 1024);
    xs[i].resize(n); ys[i].resize(n);
    int64_t w = 0;
    for (size_t j = 0; j < n; ++j) { xs[i][j] = dist(engine); ys[i][j] = dist(engine); w += (xs[i][j] - ys[i][j]) * (xs[i][j] - ys[i][j]); }
    want[i] = centre(w);
  }
  auto worker = [&](int tid) {
    OracleCiphertextFactory f(base);   // shares parameters, keys and the nonce counter
    auto astInput = Parser::parse("secret int x = {0};\nsecret int y = {0};");   // placeholders: values come from setNextInputs
    auto astProgram = Parser::parse(ladder.str());
    auto astOutput = Parser::parse("s = s;");
    TypeCheckingVisitor tcv;
    auto rootScope = std::make_unique<Scope>(*astProgram);
    for (const char *v : {"x", "y"}) {
      auto sid = std::make_unique<ScopedIdentifier>(*rootScope, v);
      rootScope->addIdentifier(v);
      tcv.addVariableDatatype(*sid, Datatype(Type::INT, true));
    }
    tcv.setRootScope(std::move(rootScope));
    astProgram->accept(tcv);
    auto tainted = tcv.getSecretTaintedNodes();
    for (int i = tid; i < total; i += T) {
      f.setNextInputs({xs[i], ys[i]});
      RuntimeVisitor srv(f, *astInput, tainted);
      srv.executeAst(*astProgram);
      auto output = srv.getOutput(*astOutput);
      for (const auto &[identifier, value] : output)
        if (auto c = dynamic_cast<AbstractCiphertext *>(value.get())) {
          std::vector<int64_t> slots;
          f.decryptCiphertext(*c, slots);
          got[i] = slots[0];
        }
    }
  };
  auto run_step = [&] {
    std::vector<pid_t> kids;
    for (int tid = 0; tid < T; ++tid) {
      pid_t p = fork();
      if (p == 0) { worker(tid); _exit(0); }
      kids.push_back(p);
    }
    for (pid_t p : kids) { int st = 0; waitpid(p, &st, 0); }
  };
  std::streambuf *old = std::cout.rdbuf(nullptr);   // the interpreter prints "Program reached return statement.." per walk
  for (int w = 0; w < warmup; ++w) run_step();
  auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < steps; ++s) run_step();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::cout.rdbuf(old);
  bool ok = true;
  for (int i = 0; i < total; ++i) ok = ok && got[i] == want[i];
  printf("{\"ops_per_s\": %.3f, \"instances_per_step\": %d, \"threads\": %d, \"steps\": %d, \"ms_per_step\": %.3f, \"result_check\": \"%s\", "
         "\"path\": \"reference Parser + TypeCheckingVisitor + RuntimeVisitor (libabc_ref.a) -> OracleCiphertextFactory -> liboracle_bfv.so\"}\n",
         (double)total * (1 + rotations) * steps / secs, total, T, steps, secs / steps * 1e3, ok ? "ok" : "FAILED");
  return ok ? 0 : 1;
}
