// emu_stubs.cc -- TEST INFRASTRUCTURE: the place for stand-ins of units that cannot be built for the CPU emulation
// (tests/emu/build_emu_lib.py).  Empty now: the slab3 and the staged cell kernels are built from their own sources, their PTX helpers
// (bulk-async copy + mbarrier, cp.async) with host bodies that copy at issue.
namespace mfg {
}
