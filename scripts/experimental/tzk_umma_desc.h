#include "../../torcheasyrec_b200/csrc/tzk_umma_desc.h"
