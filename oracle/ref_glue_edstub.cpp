// TEST INFRASTRUCTURE -- the EDLines detector (extractor: 1, not the shipped configuration) lives in ED_Lib, which is
// not compiled; LSDDetector_custom.cpp only references these three entry points.  They abort if ever reached.
// (Same stubs as at the end of ref_glue.cpp, for the libraries that do not link ref_glue.o.)
#include <cstdio>
#include <cstdlib>
#define PLVIREF_STUB(fn, sym)                                                         \
  extern "C" void fn() __asm__(sym);                                                  \
  void fn() { fprintf(stderr, "libplvi_ref: EDLines is not part of this build\n"); abort(); }
PLVIREF_STUB(plviref_stub_edlines_ctor0, "_ZN7EDLinesC1Ev")
PLVIREF_STUB(plviref_stub_edlines_ctor1, "_ZN7EDLinesC1EN2cv3MatEdidd")
PLVIREF_STUB(plviref_stub_edlines_getlines, "_ZN7EDLines8getLinesEv")
