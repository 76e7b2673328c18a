/* types.h -- element and array types of the drop-in C API.
 *
 * Replaces reference include/types.h:4-6 (`#define DataType float`, `typedef DataType *Array`).
 * The legacy-named entry points (inverse_gpu.h, gauss_gpu.h, helper_cpu.h) are single
 * precision exactly like the reference; double precision is reached through the
 * `_f64` entry points of invgpu.h, so no second compile of the library is needed.
 */
#ifndef INVGPU_TYPES_H
#define INVGPU_TYPES_H

#ifndef DataType
#define DataType float
#endif

typedef DataType *Array;

#endif /* INVGPU_TYPES_H */
