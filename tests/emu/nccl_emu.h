// nccl_emu.h -- TEST INFRASTRUCTURE ONLY: the handful of NCCL declarations dist.cu uses, backed by
// an in-process stand-in (nccl_emu.cpp) in which every rank is an OS thread of the test process.
#pragma once
#include <cstddef>

#include "cuda_emu.h"

typedef enum { ncclSuccess = 0, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef struct emuNcclComm* ncclComm_t;
