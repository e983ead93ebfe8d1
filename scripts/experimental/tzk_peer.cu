// stand-alone build of torcheasyrec_b200/csrc/tzk_peer.cu (host shim tests)
#include "../../torcheasyrec_b200/csrc/tzk_peer.cu"
