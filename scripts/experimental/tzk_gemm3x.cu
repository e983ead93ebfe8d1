// stand-alone build of torcheasyrec_b200/csrc/tzk_gemm3x.cu (host emulation tests, try_gemm3x.py)
#include "../../torcheasyrec_b200/csrc/tzk_gemm3x.cu"
