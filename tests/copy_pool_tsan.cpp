// ThreadSanitizer build of the host copy pool (tests/test_boundary_cpu.py): many back-to-back jobs of varying size.
#include "../aerobulk_b200/csrc/ab_copy_pool.hpp"
#include <cstdio>
#include <vector>
int main(int argc, char **)
{
    abpool::CopyPool &pool = *new abpool::CopyPool(6, argc > 1);   // never destroyed, like the library's (detached threads)
    const size_t N = 1 << 18;
    std::vector<double> a(N), b(N);
    int bad = 0;
    for (int r = 0; r < 400; ++r) {
        const size_t n = 1 + (size_t)((r * 7919u) % N), piece = 1 + (size_t)((r * 131u) % 5000);
        for (size_t i = 0; i < n; ++i) a[i] = (double)(i + r);
        std::vector<abpool::CopyPiece> jobs;
        for (size_t o = 0; o < n; o += piece) jobs.push_back({b.data() + o, a.data() + o + 0, sizeof(double) * (n - o < piece ? n - o : piece)});
        pool.run(jobs.data(), (int)jobs.size());
        bad += memcmp(a.data(), b.data(), sizeof(double) * n) != 0;
        if (r % 50 == 49) std::this_thread::sleep_for(std::chrono::milliseconds(3));   // let the workers fall asleep
    }
    printf("bad=%d\n", bad);
    return bad != 0;
}
