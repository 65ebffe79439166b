// Instantiation groups of the FFMA path kernel.  Compiled once per group
// (-DMCR_INST_GROUP=g) so the fully unrolled kernels build in parallel; each group owns its
// own copy of the __constant__ transform storage.
#define MCR_FFMA_MAX_N_INTERNAL 128
#ifndef MCR_PPT2_MAX_NT
#define MCR_PPT2_MAX_NT 0
#endif
#ifndef MCR_DEV_PPT
#define MCR_DEV_PPT 2
#endif
#include "mcr_ffma.cuh"

#ifndef MCR_INST_GROUP
#error "compile with -DMCR_INST_GROUP=<0..7>"
#endif

namespace mcr {

// trials per thread: 2 while both accumulator sets fit the register file without spilling
template <int NT> constexpr int ppt_for() { return NT <= MCR_PPT2_MAX_NT ? 2 : 1; }

template <int NT, bool DENSE_TOO>
static void reg_size() {
    register_variant(make_variant<NT, true, false, ppt_for<NT>()>());
    register_variant(make_variant<NT, true, true, 1>());
    if constexpr (DENSE_TOO) {
        register_variant(make_variant<NT, false, false, ppt_for<NT>()>());
        register_variant(make_variant<NT, false, true, 1>());
    }
}

#define MCR_GROUP_FN_(g) register_group##g
#define MCR_GROUP_FN(g) MCR_GROUP_FN_(g)

void MCR_GROUP_FN(MCR_INST_GROUP)() {
#if defined(MCR_DEV_SINGLE)
    // development build: one size only (fast rebuilds for kernel tuning)
#if MCR_INST_GROUP == 1
    register_variant(make_variant<MCR_DEV_SINGLE, true, false, MCR_DEV_PPT>());
#endif
#elif MCR_INST_GROUP == 0
    reg_size<8, true>(); reg_size<16, true>(); reg_size<24, true>(); reg_size<32, true>();
#elif MCR_INST_GROUP == 1
    reg_size<40, true>(); reg_size<48, false>(); reg_size<50, true>();
#elif MCR_INST_GROUP == 2
    reg_size<56, false>(); reg_size<64, true>();
#elif MCR_INST_GROUP == 3
    reg_size<72, false>(); reg_size<80, true>();
#elif MCR_INST_GROUP == 4
    reg_size<88, false>(); reg_size<96, false>();
#elif MCR_INST_GROUP == 5
    reg_size<100, true>();
#elif MCR_INST_GROUP == 6
    reg_size<112, false>();
#elif MCR_INST_GROUP == 7
    reg_size<128, false>();
#endif
}

}  // namespace mcr
