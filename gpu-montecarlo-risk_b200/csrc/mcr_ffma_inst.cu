// Instantiation groups of the FFMA path kernel.  Compiled once per group
// (-DMCR_INST_GROUP=g) so the fully unrolled kernels build in parallel; each group owns its
// own copy of the __constant__ transform storage.
#define MCR_FFMA_MAX_N_INTERNAL 128
#include "mcr_ffma.cuh"

#ifndef MCR_INST_GROUP
#error "compile with -DMCR_INST_GROUP=<0..7>"
#endif

namespace mcr {

template <int NT, bool DENSE_TOO>
static void reg_size() {
    register_variant(make_variant<NT, true, false>());
    register_variant(make_variant<NT, true, true>());
    if constexpr (DENSE_TOO) {
        register_variant(make_variant<NT, false, false>());
        register_variant(make_variant<NT, false, true>());
    }
}

#define MCR_GROUP_FN_(g) register_group##g
#define MCR_GROUP_FN(g) MCR_GROUP_FN_(g)

void MCR_GROUP_FN(MCR_INST_GROUP)() {
#if MCR_INST_GROUP == 0
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
