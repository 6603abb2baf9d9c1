// NVTX3 ranges in the reference's "dtFFT" domain, with its range names and ARGB colours
// (PHASE_BEGIN / REGION_BEGIN, src/include/_dtfft_profile.h:1-19; domain and push / pop
// src/interfaces/external/dtfft_interface_nvtx.F90:60-100, dtfft_interface_nvtx3.c:3-24; colours
// src/dtfft_parameters.F90:317-405).  The reference compiles them in with -DDTFFT_WITH_PROFILER; NVTX3
// is header-only and a few nanoseconds per call when no tool is attached, so they are always on here
// (DTFFTB_NVTX=0 switches them off).  Visible in `nsys` timelines and `ncu --nvtx`.
#pragma once
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdlib>

namespace dtfftb {

constexpr uint32_t kColorCreate = 0x00FAB53C, kColorExecute = 0x00E25DFC, kColorTranspose = 0x00B175BD,
                   kColorFft = 0x00FCD05D, kColorAutotune = 0x006075FF, kColorDestroy = 0x00000000;
// COLOR_TRANSPOSE_PALLETTE(-3:3), COLOR_RESHAPE_PALLETTE(11:14)
constexpr uint32_t kColorTransposeType[7] = {0x007A6D7D, 0x008C826A, 0x0076A797, 0, 0x005DFCCA, 0x00E3CF9F, 0x00546F66};
constexpr uint32_t kColorReshapeType[4] = {0x0000FF00, 0x00FF00FF, 0x004B0082, 0x00CD853F};

class TraceRange {
public:
    TraceRange(const char* name, uint32_t argb) {
        if (!enabled()) return;
        nvtxEventAttributes_t a = {};
        a.version = NVTX_VERSION;
        a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
        a.messageType = NVTX_MESSAGE_TYPE_ASCII;
        a.message.ascii = name;
        a.colorType = NVTX_COLOR_ARGB;
        a.color = argb;
        nvtxDomainRangePushEx(domain(), &a);
        pushed_ = true;
    }
    ~TraceRange() {
        if (pushed_) nvtxDomainRangePop(domain());
    }
    TraceRange(const TraceRange&) = delete;
    TraceRange& operator=(const TraceRange&) = delete;

private:
    static bool enabled() {
        static const bool on = [] {
            const char* e = getenv("DTFFTB_NVTX");
            return !(e && e[0] == '0');
        }();
        return on;
    }
    static nvtxDomainHandle_t domain() {
        static nvtxDomainHandle_t d = nvtxDomainCreateA("dtFFT");
        return d;
    }
    bool pushed_ = false;
};

}  // namespace dtfftb
