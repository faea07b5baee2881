// Miscellaneous C-ABI entry points: version, error strings, device check.
#include "hvlm_internal.cuh"

extern "C" const char* hvlm_strerror(int status) {
    switch (status) {
        case HVLM_OK: return "ok";
        case HVLM_ERR_BAD_ARG: return "bad argument (null pointer, negative size or unknown enum)";
        case HVLM_ERR_BAD_SHAPE: return "unsupported shape";
        case HVLM_ERR_BAD_DTYPE: return "unsupported dtype";
        case HVLM_ERR_ALIGN: return "pointer not sufficiently aligned";
        case HVLM_ERR_CUDA: return "CUDA failure (launch/driver error, or no sm_100 device)";
        case HVLM_ERR_WORKSPACE: return "workspace too small";
        case HVLM_ERR_UNSUPPORTED: return "unsupported mode";
        default: return "unknown hvlm status";
    }
}

extern "C" int hvlm_abi_version(void) { return HVLM_ABI_VERSION; }

extern "C" int hvlm_device_check(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return HVLM_ERR_CUDA;
    }
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess)
        return HVLM_ERR_CUDA;
    // the cubin is sm_100a only (tcgen05 / TMEM); it does not run on any other architecture
    return (major == 10 && minor == 0) ? HVLM_OK : HVLM_ERR_CUDA;
}
