// Internal launch interfaces shared by engine.cu, conv_ref.cu and conv_umma.cu.
#pragma once
#include "common.cuh"

namespace fnnu {

struct ConvArgs {
  const __half* src;        // first input channel of sample 0
  int src_cs;
  const double* src_stats;  // sums of the source channels (offset to the first input channel)
  const ChanMeta* src_meta;
  int src_stat_stride;
  double src_inv_count;
  __half* dst;              // first output channel of sample 0
  int dst_cs;
  double* dst_stats;        // or nullptr (no InstanceNorm after this op)
  int dst_stat_stride;
  const float* w;           // direct kernel: [tap][cin][cout_pad] fp32
  const void* w_umma;       // tcgen05 kernels: packed fp16 blob (conv_umma.cu / conv_umma_rows.cu) or nullptr
  int use_rows;             // w_umma is packed for a row-streaming kernel: 1 = conv_umma_rows.cu, 2 = conv_umma_zrows.cu
  const float* bias;        // [cout] or nullptr
  int cin, cout, cout_pad;
  int in_d[3], out_d[3];
  int k[3], s[3], pad[3];
  int ntaps;
  int transposed;
  int batch;
};

struct EltArgs {
  const __half* src;
  int src_cs;
  const double* src_stats;
  const ChanMeta* src_meta;
  int src_stat_stride;
  const __half* src2;
  int src2_cs;
  const double* src2_stats;
  const ChanMeta* src2_meta;
  int src2_stat_stride;
  __half* dst;
  int dst_cs;
  int d[3];     // output dims
  int s[3];     // pooling stride (avgpool)
  int c;
  int batch;
  double inv_count;   // 1 / voxels of the SOURCE buffer(s)
  float slope;
};

int num_sms();

int launch_conv_direct(const ConvArgs& a, cudaStream_t s);
// CUDA-core specialisations: first layer (Cin <= 4) and the 1x1x1 segmentation head (<= 8 heads)
bool direct_specialised(const ConvArgs& a);
bool prefer_cuda_cores(const ConvArgs& a);     // true where the specialised kernel beats the tensor-core path
int launch_conv_specialised(const ConvArgs& a, cudaStream_t s);
int launch_pack_weights_direct(const float* w_dev, float* out, int cin, int cout, int cout_pad, int ntaps,
                               int transposed, cudaStream_t s);
int launch_add_act(const EltArgs& a, cudaStream_t s);
int launch_avgpool(const EltArgs& a, cudaStream_t s);

// tcgen05 implicit-GEMM back end (conv_umma.cu)
bool umma_supported(const ConvArgs& a);
size_t umma_packed_weight_bytes(int cin, int cout, int ntaps, int transposed);
int launch_pack_weights_umma(const float* w_dev, void* out, const ConvArgs& a, cudaStream_t s);
int launch_conv_umma(const ConvArgs& a, cudaStream_t s);

// row-streaming, ky-folded tcgen05 kernel for thin full-resolution layers (conv_umma_rows.cu)
bool rows_supported(const ConvArgs& a);
int launch_pack_weights_rows(const float* w_dev, void* out, const ConvArgs& a, cudaStream_t s);
int launch_conv_rows(const ConvArgs& a, cudaStream_t s);

// row-streaming kernel, second generation (conv_umma_zrows.cu): 3x3x3 / 1x3x3, stride 1, Cin 16 | 32, Cout <= 32,
// 64 <= W <= 128; two output planes per pass when Cout <= 16, 3x3x3 and the depth is even; FNNU_ZROWS=0 disables it
bool zrows_supported(const ConvArgs& a);
int launch_pack_weights_zrows(const float* w_dev, void* out, const ConvArgs& a, cudaStream_t s);
int launch_conv_zrows(const ConvArgs& a, cudaStream_t s);

// conv_first_umma.cu: Cin == 1 first layer with the 27 taps as the K dimension of the MMA (round 1; runs only when
// conv_first_zpair.cu is switched off or does not take the shape; FNNU_FIRST_LAYER_TC=0 selects the CUDA-core kernel)
bool first_umma_supported(const ConvArgs& a);
int launch_conv_first_umma(const ConvArgs& a, cudaStream_t s);
// conv_first_zpair.cu: the same layer with (plane, ky) as K, kx as a descriptor shift and two output planes per MMA
// (Cout padded to 16 or 32; FNNU_FIRST_ZPAIR=0 falls back to conv_first_umma.cu)
bool first_zpair_supported(const ConvArgs& a);
int launch_conv_first_zpair(const ConvArgs& a, cudaStream_t s);

// conv_tconv_umma.cu: transposed conv with kernel == stride as one GEMM per 128 input voxels (Cin 32 / 64 / 128,
// taps * Cout_pad <= 512; FNNU_TCONV_UMMA=0 leaves the layer to the generic kernel)
bool tconv_umma_supported(const ConvArgs& a);
int launch_tconv_umma(const ConvArgs& a, cudaStream_t s);

// conv_s2_umma.cu: stride-2 3x3x3 conv 16 -> <= 32 channels, row-streaming over the NDHWC row seen as position pairs
// (FNNU_S2_UMMA=0 leaves the layer to the generic kernel)
bool s2_umma_supported(const ConvArgs& a);
int launch_conv_s2_umma(const ConvArgs& a, cudaStream_t s);

}  // namespace fnnu
