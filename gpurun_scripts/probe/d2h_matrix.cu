// d2h_matrix.cu — device-to-host result copies when several ranks share one host: per-rank and
// aggregate GB/s for R ranks x K concurrent copy streams x host memory kind.  One process per rank
// (fork), started together through a pipe barrier, like the ranks of bench.py.
//   usage: d2h_matrix <max_ranks> [bytes_per_rank_MiB]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

enum { PINNED = 0, REGISTERED_HUGE = 1, WRITE_COMBINED = 2 };
static const char *kind_name[] = {"cudaHostAlloc", "mmap(2MiB huge)+cudaHostRegister", "cudaHostAllocWriteCombined"};

static void *host_buf(int kind, size_t bytes)
{
    void *p = nullptr;
    if (kind == PINNED) { if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr; }
    else if (kind == WRITE_COMBINED) { if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) return nullptr; }
    else {
        p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
        if (p == MAP_FAILED) {           // no huge pages reserved: transparent huge pages instead
            p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (p == MAP_FAILED) return nullptr;
            madvise(p, bytes, MADV_HUGEPAGE);
        }
        memset(p, 1, bytes);
        if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) return nullptr;
    }
    return p;
}

int main(int argc, char **argv)
{
    const int max_ranks = argc > 1 ? atoi(argv[1]) : 1;
    const size_t bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 2048) << 20;
    int ndev = 0;
    // (no CUDA call in the parent before fork)
    for (int kind = 0; kind < 3; kind++)
        for (int ranks = 1; ranks <= max_ranks; ranks *= 2)
            for (int k = 1; k <= 8; k *= 2) {
                int go[2], res[2];
                if (pipe(go) || pipe(res)) return 1;
                for (int r = 0; r < ranks; r++) {
                    if (fork() == 0) {
                        cudaGetDeviceCount(&ndev);
                        cudaSetDevice(r % (ndev ? ndev : 1));
                        void *d = nullptr;
                        cudaMalloc(&d, bytes);
                        cudaMemset(d, 7, bytes);
                        void *h = host_buf(kind, bytes);
                        cudaStream_t s[8];
                        for (int i = 0; i < k; i++) cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking);
                        double best = 0;
                        char c;
                        if (h) {
                            const size_t piece = bytes / k;
                            for (int rep = 0; rep < 4; rep++) {
                                if (rep == 1) { c = 'r'; (void)!write(res[1], &c, 1); (void)!read(go[0], &c, 1); }   // ready -> go
                                cudaDeviceSynchronize();
                                const double t0 = now();
                                for (int i = 0; i < k; i++)
                                    cudaMemcpyAsync((char *)h + i * piece, (char *)d + i * piece, piece, cudaMemcpyDeviceToHost, s[i]);
                                cudaDeviceSynchronize();
                                const double gbs = bytes / (now() - t0) / 1e9;
                                if (rep >= 1 && (best == 0 || gbs < best)) best = gbs;      // slowest timed pass, like max-over-ranks timing
                            }
                        } else { c = 'r'; (void)!write(res[1], &c, 1); (void)!read(go[0], &c, 1); }
                        char line[64];
                        int n = snprintf(line, sizeof line, "%.2f\n", best);
                        (void)!write(res[1], line, n);
                        _exit(0);
                    }
                }
                char c;
                for (int r = 0; r < ranks; r++) (void)!read(res[0], &c, 1);      // all ranks are ready
                for (int r = 0; r < ranks; r++) (void)!write(go[1], "g", 1);
                while (wait(nullptr) > 0) {}
                char buf[1024];
                const int n = (int)read(res[0], buf, sizeof buf - 1);
                buf[n > 0 ? n : 0] = 0;
                double sum = 0, mn = 1e9;
                for (char *t = strtok(buf, "\n"); t; t = strtok(nullptr, "\n")) { const double v = atof(t); sum += v; if (v < mn) mn = v; }
                printf("%-34s ranks %d copies/rank %d : slowest rank %6.2f GB/s  sum %7.2f GB/s  ranks x slowest %7.2f GB/s\n",
                       kind_name[kind], ranks, k, mn, sum, ranks * mn);
                fflush(stdout);
                close(go[0]); close(go[1]); close(res[0]); close(res[1]);
            }
    return 0;
}
