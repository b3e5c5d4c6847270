// Runs the REAL fp16 + e4m3 operand conversion (split16_f16f8) and weight packers (pack_b_f16f8_elem,
// pack_conv_f16f8_elem) of csrc/dce_tc.cuh ON THE HOST — they are __host__ __device__ — and writes their bytes to a
// file, so tests/test_f16f8_models_cpu.py can compare the CUDA source byte for byte with the numpy emulation
// (tools/emulate_f16f8.py) whose layouts the emulated MMAs consume.  No GPU, no CUDA call.
//   nvcc -std=c++17 -O1 -o /tmp/host_check_f16f8 tools/host_check_f16f8.cu && /tmp/host_check_f16f8 in.bin out.bin
// in.bin : int32 header {n_split, signed, fc_n, fc_k, fc_bn, fc_kind, conv_cout, conv_cin, conv_cin_pad}, float sw_fc, float sw_conv,
//          then n_split*16 floats, fc_n*fc_src_k floats (W of the Linear layer), conv_cout*conv_cin*3 floats
// out.bin: per split: 32 B fp16, 16 B lo8, 16 B hi8; then the Linear image; then the conv image
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../deep_contact_estimator_b200/csrc/dce_tc.cuh"

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int hdr[9];
    float sw[2];
    if (fread(hdr, 4, 9, f) != 9 || fread(sw, 4, 2, f) != 2) return 4;
    const int n_split = hdr[0], is_signed = hdr[1], fc_n = hdr[2], fc_k = hdr[3], fc_bn = hdr[4], fc_kind = hdr[5];
    const int cout = hdr[6], cin = hdr[7], cin_pad = hdr[8];
    const size_t fc_src = (size_t)fc_n * (fc_kind == 3 ? 4736 : fc_k);
    std::vector<float> y((size_t)n_split * 16), wfc(fc_src), wcv((size_t)cout * cin * 3);
    if (fread(y.data(), 4, y.size(), f) != y.size() || fread(wfc.data(), 4, wfc.size(), f) != wfc.size() ||
        fread(wcv.data(), 4, wcv.size(), f) != wcv.size()) return 5;
    fclose(f);
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 6;
    for (int i = 0; i < n_split; ++i) {
        uint4 fa, fb, lo8, hi8;
        if (is_signed) dce::tc::split16_f16f8<true>(y.data() + 16 * i, fa, fb, lo8, hi8);
        else dce::tc::split16_f16f8<false>(y.data() + 16 * i, fa, fb, lo8, hi8);
        fwrite(&fa, 16, 1, o); fwrite(&fb, 16, 1, o); fwrite(&lo8, 16, 1, o); fwrite(&hi8, 16, 1, o);
    }
    {
        const int n_tiles = fc_n / fc_bn, stages = fc_k / 32;
        std::vector<uint8_t> img((size_t)n_tiles * stages * 8 * fc_bn * 16, 0);
        for (size_t idx = 0; idx < (size_t)fc_n * fc_k; ++idx)
            dce::tc::pack_b_f16f8_elem(wfc.data(), img.data(), idx, stages, fc_bn, fc_kind, fc_k, sw[0]);
        fwrite(img.data(), 1, img.size(), o);
    }
    {
        std::vector<uint8_t> img((size_t)2 * (cin_pad / 32) * 192 * cout, 0);
        for (int idx = 0; idx < cout * cin_pad * 3; ++idx)
            dce::tc::pack_conv_f16f8_elem(wcv.data(), img.data(), idx, cout, cin, cin_pad, sw[1]);
        fwrite(img.data(), 1, img.size(), o);
    }
    // layout helper tables (all as uint64): for C in {64, 128, 2048, 4736}: per 16-channel group g16 the tape
    // destinations (part_stride = 1 << 40, kch_stride = 1 << 20, so chunk indices can be read off) and the slab
    // destinations; then f8_stage_src for fc.0 (148 stages) and fc.3 (64 stages), [s][part][j]
    const int Cs[4] = {64, 128, 2048, 4736};
    for (int ci = 0; ci < 4; ++ci)
        for (int g = 0; g < Cs[ci] / 16; ++g) {
            const dce::tc::F8Dst t = dce::tc::f8_tape_dst(g, Cs[ci], (size_t)1 << 40, (size_t)1 << 20);
            const dce::tc::F8Dst sl = dce::tc::f8_slab_dst(g, Cs[ci]);
            const unsigned long long v[6] = {t.f16, t.lo8, t.hi8, sl.f16, sl.lo8, sl.hi8};
            fwrite(v, 8, 6, o);
        }
    const int St[2] = {148, 64};
    for (int li = 0; li < 2; ++li)
        for (int s = 0; s < St[li]; ++s)
            for (int part = 0; part < 2; ++part)
                for (int j = 0; j < 4; ++j) {
                    const unsigned long long v = dce::tc::f8_stage_src(s, St[li], part, j, (size_t)1 << 40, (size_t)1 << 20);
                    fwrite(&v, 8, 1, o);
                }
    // B-operand offsets inside a conv weight block: [cout in {64, 128}][img][tap] e4m3, then [cout][tap][kk] fp16
    for (int cout_ : {64, 128}) {
        for (int img = 0; img < 2; ++img)
            for (int tap = 0; tap < 3; ++tap) { const unsigned long long v = dce::tc::f8_wblk_e4m3(cout_, img, tap); fwrite(&v, 8, 1, o); }
        for (int tap = 0; tap < 3; ++tap)
            for (int kk = 0; kk < 2; ++kk) { const unsigned long long v = dce::tc::f8_wblk_f16(cout_, tap, kk); fwrite(&v, 8, 1, o); }
    }
    // issue plans, 4 x uint64 per MMA {a_off, b_off, e4m3, mode}: f8_fc_mma for fc.0 (148 stages, BN 256) and fc.3 (64, BN 128)
    // as [s][i]; f8_conv_mma for block1 (half 2, C 64, cout 64), conv3 (2, 64, 128), conv4 (4, 128, 128) as [s][tap][i]
    const int fc_cfg[2][2] = {{148, 256}, {64, 128}};
    for (auto& c : fc_cfg)
        for (int s = 0; s < c[0]; ++s)
            for (int i = 0; i < 4; ++i) {
                const dce::tc::F8Mma m = dce::tc::f8_fc_mma(s, c[0], i, 4 * 130 * 16, 4 * c[1] * 16, c[1] * 16);
                const unsigned long long v[4] = {m.a_off, m.b_off, m.e4m3, m.mode};
                fwrite(v, 8, 4, o);
            }
    const int cv_cfg[3][3] = {{2, 64, 64}, {2, 64, 128}, {4, 128, 128}};
    for (auto& c : cv_cfg)
        for (int s = 0; s < 2 * c[0]; ++s)
            for (int tap = 0; tap < 3; ++tap)
                for (int i = 0; i < 2; ++i) {
                    const dce::tc::F8Mma m = dce::tc::f8_conv_mma(s, c[0], tap, i, c[1], c[2]);
                    const unsigned long long v[4] = {m.a_off, m.b_off, m.e4m3, m.mode};
                    fwrite(v, 8, 4, o);
                }
    fclose(o);
    return 0;
}
