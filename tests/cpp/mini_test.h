// tests/cpp/mini_test.h -- a few lines standing in for the Catch macros the reference's tests use
// (TEST_CASE / REQUIRE), so the restated tests read like reference tests/test_cotan.cpp / test_trajectory.cpp.
#ifndef MINI_TEST_H
#define MINI_TEST_H
#include <cstdio>
#include <cstdlib>
#include <vector>
struct MiniTestCase { const char *name; void (*fn)(); };
inline std::vector<MiniTestCase> &miniTests() { static std::vector<MiniTestCase> t; return t; }
struct MiniTestRegistrar { MiniTestRegistrar(const char *n, void (*f)()) { MiniTestCase c = {n, f}; miniTests().push_back(c); } };
static int g_miniFailures = 0;
#define MINI_CAT2(a, b) a##b
#define MINI_CAT(a, b) MINI_CAT2(a, b)
#define TEST_CASE(name) static void MINI_CAT(miniTest, __LINE__)(); static MiniTestRegistrar MINI_CAT(miniReg, __LINE__)(name, MINI_CAT(miniTest, __LINE__)); static void MINI_CAT(miniTest, __LINE__)()
#define REQUIRE(expr) do { if (!(expr)) { std::printf("FAILED %s:%d: REQUIRE(%s)\n", __FILE__, __LINE__, #expr); ++g_miniFailures; } } while (0)
#define MINI_TEST_MAIN int main() { for (size_t i = 0; i < miniTests().size(); ++i) { std::printf("[ RUN ] %s\n", miniTests()[i].name); miniTests()[i].fn(); } std::printf("%s (%d failure(s))\n", g_miniFailures ? "FAILED" : "ALL PASSED", g_miniFailures); return g_miniFailures ? 1 : 0; }
#endif
