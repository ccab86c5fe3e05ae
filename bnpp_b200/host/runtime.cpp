#include "runtime.hh"

#include <cstdio>
#include <cstdlib>

namespace bn {
namespace gpu {

static bnpp_ctx *g_ctx = nullptr;

bnpp_ctx *ctx()
{
    if (!g_ctx) {
        // load every kernel when the context is created instead of at its first launch: the
        // reference's `uptime` brackets only the inference call, and so must ours
        setenv("CUDA_MODULE_LOADING", "EAGER", 0);
        const char *dev = std::getenv("BNPP_DEVICE");
        const int rc = bnpp_ctx_create(dev ? std::atoi(dev) : 0, nullptr, &g_ctx);
        if (rc != BNPP_OK) {
            std::fprintf(stderr, "bn-pp (B200 build): cannot create a CUDA context: %s\n", bnpp_last_error(nullptr));
            std::exit(3);
        }
        // grow the stream-ordered memory pool now (its release threshold is unlimited, bnpp_ctx_create): the first large
        // allocation of a query -- the arena of the plan's intermediates -- otherwise pays for the growth, 1-500 ms
        // depending on the box, inside the timed inference call.  BNPP_POOL_MB sets the size (default 1024, 0 = off).
        const char *pm = std::getenv("BNPP_POOL_MB");
        const unsigned long mb = pm ? std::strtoul(pm, nullptr, 10) : 1024ul;
        if (mb) {
            double *warm = nullptr;
            if (bnpp_alloc(g_ctx, (uint64_t)mb * (1ull << 20) / 8, &warm) == BNPP_OK) {
                bnpp_free(g_ctx, warm);
                bnpp_ctx_sync(g_ctx);
            }
        }
        std::atexit(shutdown);
    }
    return g_ctx;
}

void check(int rc, const char *what)
{
    if (rc == BNPP_OK) return;
    std::fprintf(stderr, "bn-pp (B200 build): %s failed (%d): %s\n", what, rc, bnpp_last_error(g_ctx));
    std::exit(4);
}

void shutdown()
{
    if (g_ctx) bnpp_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

}  // namespace gpu
}  // namespace bn
