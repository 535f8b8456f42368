// tests/emu/emu_runtime.cpp — TEST INFRASTRUCTURE ONLY: warp-as-fibers scheduler (see cuda_runtime.h).
#include <cuda_runtime.h>

#include <vector>

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
namespace {
constexpr int WARP = 32;
constexpr size_t STACK = 256 * 1024;
ucontext_t g_sched, g_lane[WARP];
std::vector<char> g_stack[WARP];
bool g_done[WARP];
unsigned g_round[WARP];
uint32_t g_xbuf[2][WARP];
int g_cur = 0;
unsigned g_tid[WARP];
const std::function<void()>* g_body = nullptr;

void trampoline() {
    (*g_body)();
    g_done[g_cur] = true;
    swapcontext(&g_lane[g_cur], &g_sched);
}
}  // namespace

int lane_id() { return g_cur; }

unsigned char* smem() {
    alignas(128) static unsigned char buf[256 * 1024];
    return buf;
}

uint32_t shfl_exchange(uint32_t v, int src) {
    int me = g_cur;
    unsigned r = g_round[me]++;
    g_xbuf[r & 1][me] = v;
    swapcontext(&g_lane[me], &g_sched);  // resumed after every live lane has written round r
    return g_xbuf[r & 1][src];
}

void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    g_body = &body;
    gridDim.x = grid; blockDim.x = block;
    for (int l = 0; l < WARP; l++) if (g_stack[l].empty()) g_stack[l].resize(STACK);
    for (unsigned b = 0; b < grid; b++) {
        for (unsigned w = 0; w * WARP < block; w++) {
            for (int l = 0; l < WARP; l++) {
                g_done[l] = (w * WARP + l >= block);
                g_round[l] = 0;
                g_tid[l] = w * WARP + l;
                if (g_done[l]) continue;
                getcontext(&g_lane[l]);
                g_lane[l].uc_stack.ss_sp = g_stack[l].data();
                g_lane[l].uc_stack.ss_size = STACK;
                g_lane[l].uc_link = &g_sched;
                makecontext(&g_lane[l], trampoline, 0);
            }
            for (;;) {
                bool any = false;
                for (int l = 0; l < WARP; l++) {
                    if (g_done[l]) continue;
                    any = true;
                    g_cur = l;
                    blockIdx.x = b; threadIdx.x = g_tid[l];
                    swapcontext(&g_sched, &g_lane[l]);
                }
                if (!any) break;
            }
        }
    }
}
}  // namespace emu
