/*
 * emba_synth -- GPU event simulator used ONLY to produce synthetic benchmark/test input
 * (SURVEY.md section 8(d)); it is not part of the drop-in boundary of include/emba_b200.h.
 * One thread per sensor pixel runs the threshold-crossing model over all simulation steps
 * (two passes: count, fill), then the events are sorted by time.
 */
#ifndef EMBA_SYNTH_H_
#define EMBA_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#define EMBA_SYNTH_API __attribute__((visibility("default")))
#else
#define EMBA_SYNTH_API
#endif
typedef struct emba_synth_s* emba_synth_t;
/* lut [Hs*Ws*3] bearing vectors, L [Hp*Wp] log-intensity panorama, R_steps [(n_steps+1)*9] row-major rotations at
 * t_start + k*dt_sim. Returns 0 on success; *n_events is the number of simulated events. */
EMBA_SYNTH_API int emba_synth_simulate(int device, int Ws, int Hs, const double* lut, int Wp, int Hp,
                                       const double* L, double C_th, int n_steps, const double* R_steps,
                                       double t_start, double dt_sim, emba_synth_t* out, int64_t* n_events);
/* time-sorted events (ties keep sensor-pixel order); arrays of *n_events entries */
EMBA_SYNTH_API int emba_synth_fetch(emba_synth_t s, uint16_t* x, uint16_t* y, int64_t* t_ns, uint8_t* pol);
EMBA_SYNTH_API int emba_synth_free(emba_synth_t s);
#ifdef __cplusplus
}
#endif
#endif
