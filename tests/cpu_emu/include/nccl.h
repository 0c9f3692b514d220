/* TEST INFRASTRUCTURE ONLY (tests/cpu_emu): the few NCCL types gf2b200.cu names.
 * The emulated library never creates an NCCL context. */
#pragma once
#include <cuda_runtime.h>
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef struct emu_nccl_comm_ *ncclComm_t;
typedef enum { ncclUint8 = 1 } ncclDataType_t;
