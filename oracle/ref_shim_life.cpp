// ORACLE (test infrastructure): C shim around the reference's checked-in generated class
// examples-old/Life-exampled/dist/Life.hpp (128x128, Open boundary, margin 1, R-pentomino init).
#include "Life.hpp"
extern "C" {
void* ref_new() { return new Life(); }
void ref_delete(void* p) { delete (Life*)p; }
void ref_init(void* p) { ((Life*)p)->init(); }
void ref_proceed(void* p) { ((Life*)p)->proceed(); }
int* ref_cell(void* p) { return ((Life*)p)->static_2_cell.data(); }
int ref_population(void* p) { return ((Life*)p)->static_0_population; }
int ref_generation(void* p) { return ((Life*)p)->static_1_generation; }
int ref_memory_size(void* p, int k) { Life* s = (Life*)p; return k == 0 ? s->memorySize0() : k == 1 ? s->memorySize1() : s->memorySize(); }
int ref_size(void* p, int k) { Life* s = (Life*)p; return k == 0 ? s->size0() : k == 1 ? s->size1() : s->size(); }
int ref_lower_margin(void* p, int k) { Life* s = (Life*)p; return k == 0 ? s->lowerMargin0() : s->lowerMargin1(); }
}
