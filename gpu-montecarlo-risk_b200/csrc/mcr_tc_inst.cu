// One (NOUT, KB) instantiation of tc_path_kernel per translation unit: -DMCR_TC_INST=i (see mcr_tc_path.cuh).
#include "mcr_tc_path.cuh"

#ifndef MCR_TC_INST
#error "compile with -DMCR_TC_INST=<variant index>"
#endif

namespace mcr {

#if MCR_TC_INST == 10   // = TC_SUPPLIED_INST: the kernel fed with the caller's normals
static_assert(MCR_TC_INST == TC_SUPPLIED_INST, "supplied-normals instantiation index");
template <>
TcVariant tc_variant_inst<MCR_TC_INST>() {
    return tc_variant<TC_SUPPLIED_NOUT, TC_SUPPLIED_KB, tc_groups(TC_SUPPLIED_NOUT, TC_SUPPLIED_KB), true>();
}
#else
template <>
TcVariant tc_variant_inst<MCR_TC_INST>() {
    constexpr int NOUT = TC_VARIANT_NOUT[MCR_TC_INST], KB = TC_VARIANT_KB[MCR_TC_INST];
    return tc_variant<NOUT, KB, tc_groups(NOUT, KB)>();
}
#endif

}  // namespace mcr
