// ORACLE (test infrastructure): C shim around the reference's checked-in generated class
// examples-old/Hydro-exampled/dist/Hydro.hpp (float, 1024x1024, Open boundary, margin 3).
#include "Hydro.hpp"
extern "C" {
void* ref_new() { return new Hydro(); }
void ref_delete(void* p) { delete (Hydro*)p; }
void ref_init(void* p) { ((Hydro*)p)->init(); }
void ref_proceed(void* p) { ((Hydro*)p)->proceed(); }
// scalar statics in declaration order: generation(int) time cfl dR0 dR1 extent0 extent1
float* ref_scalar(void* p, int idx) {
  Hydro* s = (Hydro*)p;
  switch (idx) {
    case 1: return &s->static_1_time;
    case 2: return &s->static_2_cfl;
    case 3: return &s->static_3_dR0;
    case 4: return &s->static_4_dR1;
    case 5: return &s->static_5_extent0;
    case 6: return &s->static_6_extent1;
  }
  return 0;
}
float* ref_array(void* p, int idx) {
  Hydro* s = (Hydro*)p;
  switch (idx) {
    case 7: return s->static_7_density.data();
    case 8: return s->static_8_velocity0.data();
    case 9: return s->static_9_velocity1.data();
    case 10: return s->static_10_pressure.data();
  }
  return 0;
}
int ref_memory_size(void* p, int k) { Hydro* s = (Hydro*)p; return k == 0 ? s->memorySize0() : k == 1 ? s->memorySize1() : s->memorySize(); }
int ref_size(void* p, int k) { Hydro* s = (Hydro*)p; return k == 0 ? s->size0() : k == 1 ? s->size1() : s->size(); }
int ref_lower_margin(void* p, int k) { Hydro* s = (Hydro*)p; return k == 0 ? s->lowerMargin0() : s->lowerMargin1(); }
}
