#include "../../torcheasyrec_b200/csrc/tzk_tcgen05_ptx.h"
