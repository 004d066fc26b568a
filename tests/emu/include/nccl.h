// TEST INFRASTRUCTURE (see cuda_runtime.h in this directory): the few NCCL
// types and prototypes plb_api.cu names (it resolves the symbols with dlopen,
// and an emulated solver is never given a communicator).
#pragma once
#include <cuda_runtime.h>
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef enum { ncclChar = 0, ncclInt = 2, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMin = 3 } ncclRedOp_t;
extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId *);
ncclResult_t ncclCommInitRank(ncclComm_t *, int, ncclUniqueId, int);
ncclResult_t ncclCommDestroy(ncclComm_t);
ncclResult_t ncclSend(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t ncclRecv(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t ncclAllReduce(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
ncclResult_t ncclGroupStart(void);
ncclResult_t ncclGroupEnd(void);
const char *ncclGetErrorString(ncclResult_t);
}
