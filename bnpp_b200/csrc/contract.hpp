// A fully resolved launch of the contraction engine: kernel variant, grid and parameter block.
// Planning (axis analysis, iteration order, class selection) is done once; replaying a
// descriptor only patches the operand / output pointers -- what a VE plan does per query.
#pragma once
#include <string>

#include "common.cuh"

namespace bnpp {

constexpr int kMaxF = 24;
struct Field {
    uint32_t mask, mul, sh;
};

struct ParamsHead {
    const double *in[kMaxK];
    double *out;
    double *partials;
    unsigned int *ticket;
    double *z;
    unsigned int *status;
    uint64_t n_items;           // output entries / V
    uint32_t cx;                // cardinality of the eliminated variable (1 = none)
    uint32_t sx[kMaxK];         // operand stride of the eliminated variable
    uint32_t sl[kMaxK];         // operand stride between the V entries of an item
    uint32_t sol;               // output stride between the V entries of an item
    uint8_t cls[kMaxK];
    uint8_t out_vec;            // 16-byte store allowed
};

// Mixed-radix iteration space: digits by multiply-high division, outermost axis first.
struct ParamsMR {
    ParamsHead h;
    uint32_t R;
    FastDiv div[kMaxR];
    uint32_t so[kMaxR];         // output stride per axis (per item on the innermost axis)
    uint32_t s[kMaxK][kMaxR];   // operand stride per axis
};

// Power-of-two iteration space: off = sum_f ((item >> sh) & mask) * mul, fields merged PER
// OPERAND, so an operand laid out like the output costs one field no matter how
// scattered the other operands' axes are.
struct ParamsP2 {
    ParamsHead h;
    uint8_t nf[kMaxK + 1];      // [K] is the output
    Field f[kMaxK + 1][kMaxF];
};

// One operand whose fastest axes are slow axes of the output ("transposed" layout) is not
// gathered from global memory lane by lane: each CTA chunk first copies the tile of it that
// the chunk needs into shared memory, reading along the operand's OWN fastest axes.
//   slot -> operand offset : sum_f ((slot  >> sh) & mask) * mul   over lf   (tile load)
//   item -> slot           : sum_f ((local >> sh) & mask) * mul   over cf   (consumption)
struct StageInfo {
    int32_t sk;                 // staged operand
    uint32_t tile;              // tile entries (power of two, <= 4096)
    uint32_t slot_j, slot_x;    // slot stride of the V bit and of the eliminated variable (0 = independent)
    uint32_t swz;               // bank swizzle: phys = slot ^ ((slot >> swz) & 31); the lanes differ in slot bits >= swz
    uint8_t nlf, ncf;
    Field lf[12], cf[12];
};
struct ParamsP2S {
    ParamsP2 b;
    StageInfo st;
};

struct LaunchDesc {
    const void *fn = nullptr;
    unsigned grid = 0;
    bool p2 = false;
    bool staged = false;
    int k = 0;
    ParamsP2S p2p;              // .b is the plain power-of-two block
    ParamsMR mrp;
    // pieces of the printable name (formatted lazily: a VE query launches hundreds of these)
    const char *variant = "";
    int C = 0, V = 0, U = 0;
    bool div = false, generic = false;
    uint32_t R = 0;
    std::string name();
    ParamsHead &head() { return p2 ? p2p.b.h : mrp.h; }
};

int contract_plan(bnpp_ctx *ctx, int k, const bnpp_operand *ops, const bnpp_scope *out_scope, int64_t elim_var,
                  int divide, double *out_dev, double *z_dev, LaunchDesc *desc);
// in / out / z replace the pointers the descriptor was planned with; they must be at least as
// aligned (a descriptor planned for 32-byte aligned operands may use 256-bit loads)
int contract_launch(bnpp_ctx *ctx, LaunchDesc &desc, const double *const *in, double *out, double *z);

}  // namespace bnpp
