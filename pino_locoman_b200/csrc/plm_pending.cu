// Entry points of include/pino_locoman_b200.h whose kernels have not landed yet: they fail loudly.
#include "plm_handle.cuh"

#define PLM_PENDING(h, name)                                  \
  do {                                                        \
    (h)->error = std::string(name) + ": not implemented yet"; \
    return 100;                                               \
  } while (0)

extern "C" {
int plm_state_integrate(plm_handle* h, const double*, const double*, int32_t, double*, void*) { PLM_PENDING(h, "plm_state_integrate"); }
int plm_state_difference(plm_handle* h, const double*, const double*, int32_t, double*, void*) { PLM_PENDING(h, "plm_state_difference"); }
int plm_rnea_dyn(plm_handle* h, const double*, const double*, const double*, const double*, int32_t, double*, double*, void*) { PLM_PENDING(h, "plm_rnea_dyn"); }
int plm_aba_dyn(plm_handle* h, const double*, const double*, const double*, const double*, int32_t, double*, double*, void*) { PLM_PENDING(h, "plm_aba_dyn"); }
int plm_dyn_gaps(plm_handle* h, int32_t, const double*, const double*, const double*, const double*, int32_t, double*, double*, void*) { PLM_PENDING(h, "plm_dyn_gaps"); }
int plm_centroidal_vel_gaps(plm_handle* h, const double*, const double*, const double*, int32_t, double*, void*) { PLM_PENDING(h, "plm_centroidal_vel_gaps"); }
int plm_com_dyn(plm_handle* h, const double*, const double*, int32_t, double*, void*) { PLM_PENDING(h, "plm_com_dyn"); }
int plm_frame_vel(plm_handle* h, int32_t, int32_t, const double*, const double*, int32_t, double*, void*) { PLM_PENDING(h, "plm_frame_vel"); }
}
