// Link seam between abi.cpp/engine.cpp and the device backend.  libjfnk.so links cuda_ops.cu
// (the only product backend); the GPU-less logic tests link tests/hostsim/host_ops.cpp instead.
#pragma once
#include <string>
#include "../../include/jfnk.h"
#include "device_ops.h"

namespace jfnk {
int backend_device_ok(std::string& why);
DeviceOps* backend_make_ops(const jfnk_config& cfg, std::string& why, int& code);
int backend_unique_id(void* id128, std::string& why);
int backend_comm_init(DeviceOps* ops, const void* id128, std::string& why);
int backend_peer_memory(DeviceOps* ops);
} // namespace jfnk
