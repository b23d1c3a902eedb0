// dop_const.cuh -- the Dormand-Prince tableaux in __constant__ memory, so that every coefficient is a
// direct c[bank][offset] operand of its DFMA/DMUL instead of a pair of UMOV immediates
// (the immediates were about a tenth of the issued instructions of the first kernel version).
#pragma once
#include "dop_coeffs.h"
namespace RAPT_NS {
struct D8Tab { double C2, C3, C4, C5, C6, C7, C8, C9, C10, C11, C12, A2_1, A3_1, A3_2, A4_1, A4_3, A5_1, A5_3, A5_4, A6_1, A6_4, A6_5, A7_1, A7_4, A7_5, A7_6, A8_1, A8_4, A8_5, A8_6, A8_7, A9_1, A9_4, A9_5, A9_6, A9_7, A9_8, A10_1, A10_4, A10_5, A10_6, A10_7, A10_8, A10_9, A11_1, A11_4, A11_5, A11_6, A11_7, A11_8, A11_9, A11_10, A12_1, A12_4, A12_5, A12_6, A12_7, A12_8, A12_9, A12_10, A12_11, B1, B6, B7, B8, B9, B10, B11, B12, ER1, ER6, ER7, ER8, ER9, ER10, ER11, ER12, BHH1, BHH2, BHH3; };
struct D5Tab { double C2, C3, C4, C5, C6, A2_1, A3_1, A3_2, A4_1, A4_2, A4_3, A5_1, A5_2, A5_3, A5_4, A6_1, A6_2, A6_3, A6_4, A6_5, A7_1, A7_3, A7_4, A7_5, A7_6, E1, E3, E4, E5, E6, E7; };
static __constant__ D8Tab g_d8 = { D8_C2, D8_C3, D8_C4, D8_C5, D8_C6, D8_C7, D8_C8, D8_C9, D8_C10, D8_C11, D8_C12, D8_A2_1, D8_A3_1, D8_A3_2, D8_A4_1, D8_A4_3, D8_A5_1, D8_A5_3, D8_A5_4, D8_A6_1, D8_A6_4, D8_A6_5, D8_A7_1, D8_A7_4, D8_A7_5, D8_A7_6, D8_A8_1, D8_A8_4, D8_A8_5, D8_A8_6, D8_A8_7, D8_A9_1, D8_A9_4, D8_A9_5, D8_A9_6, D8_A9_7, D8_A9_8, D8_A10_1, D8_A10_4, D8_A10_5, D8_A10_6, D8_A10_7, D8_A10_8, D8_A10_9, D8_A11_1, D8_A11_4, D8_A11_5, D8_A11_6, D8_A11_7, D8_A11_8, D8_A11_9, D8_A11_10, D8_A12_1, D8_A12_4, D8_A12_5, D8_A12_6, D8_A12_7, D8_A12_8, D8_A12_9, D8_A12_10, D8_A12_11, D8_B1, D8_B6, D8_B7, D8_B8, D8_B9, D8_B10, D8_B11, D8_B12, D8_ER1, D8_ER6, D8_ER7, D8_ER8, D8_ER9, D8_ER10, D8_ER11, D8_ER12, D8_BHH1, D8_BHH2, D8_BHH3 };
static __constant__ D5Tab g_d5 = { D5_C2, D5_C3, D5_C4, D5_C5, D5_C6, D5_A2_1, D5_A3_1, D5_A3_2, D5_A4_1, D5_A4_2, D5_A4_3, D5_A5_1, D5_A5_2, D5_A5_3, D5_A5_4, D5_A6_1, D5_A6_2, D5_A6_3, D5_A6_4, D5_A6_5, D5_A7_1, D5_A7_3, D5_A7_4, D5_A7_5, D5_A7_6, D5_E1, D5_E3, D5_E4, D5_E5, D5_E6, D5_E7 };
}  // namespace RAPT_NS
#define T8(n) g_d8.n
#define T5(n) g_d5.n
