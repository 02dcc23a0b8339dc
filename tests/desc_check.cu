#include <cstdio>
#include <cute/arch/mma_sm100_desc.hpp>
#include "ptx.cuh"
int main() {
  using namespace cute::UMMA;
  int bad = 0;
  for (uint32_t addr : {0u, 1024u, 49152u, 0x30000u, 0x38400u}) {
    SmemDescriptor d; d.desc_ = 0;
    d.version_ = 1; d.lbo_mode_ = 0; d.layout_type_ = uint8_t(LayoutType::SWIZZLE_128B);
    d.start_address_ = uint16_t(addr >> 4); d.base_offset_ = 0;
    d.stride_byte_offset_ = 1024 >> 4; d.leading_byte_offset_ = 1;
    uint64_t mine = e2b::umma_desc_kmajor_sw128(addr);
    if (mine != d.desc_) { printf("smem desc mismatch addr %x: %llx vs %llx\n", addr, (unsigned long long)mine, (unsigned long long)d.desc_); bad++; }
  }
  for (auto mn : {std::pair<int,int>{128,256}, {128,128}, {128,64}}) {
    InstrDescriptor i; i.desc_ = 0;
    i.a_format_ = uint8_t(F16F32Format::BF16); i.b_format_ = uint8_t(F16F32Format::BF16); i.c_format_ = uint8_t(CFormat::F32);
    i.a_major_ = uint8_t(Major::K); i.b_major_ = uint8_t(Major::K);
    i.m_dim_ = mn.first >> 4; i.n_dim_ = mn.second >> 3;
    uint32_t mine = e2b::umma_idesc_bf16(mn.first, mn.second);
    if (mine != i.desc_) { printf("idesc mismatch %dx%d: %x vs %x\n", mn.first, mn.second, mine, i.desc_); bad++; }
  }
  printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
