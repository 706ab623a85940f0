# every documented environment switch still produces oracle-parity results (smoke() compares with the fp32 oracle)
mkdir -p gpurun_out
for kv in "X=1" "VITATK_GEMM_2CTA=0" "VITATK_LN_FOLD=0" "VITATK_FUSE_STATS=0" "VITATK_LN_STREAM=0" "VITATK_GEMM_EPI16=0" \
          "VITATK_TC_CONST=0" "VITATK_ZIGZAG=0" "VITATK_PDL=1" "VITATK_PDL=0" "VITATK_GELU=f32" "VITATK_FUSE_DELTA=0" \
          "VITATK_LORA_MERGE=1" "VITATK_RES_F16=0" "VITATK_QKV_PACKED=0" "VITATK_TT_SITES=63" "VITATK_LN_BT=1"; do
  out=$(env $kv timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1)
  echo "$kv -> $out"
done
