# A/B timing of GEMM epilogue experiment builds (tools/build_variant.py) on the short-K shapes of cfg2
SHAPES="${SHAPES:-gemm1536_512_0_0_0 gemm2048_512_0 gemm512_512_2 gemm512_2048_2}"
for v in ${VARIANTS:-base gnobias gnostore gnobiasnostore gnoepi}; do
  if [ $v = base ]; then unset WFL_LIB; else export WFL_LIB=$PWD/variants/libwfl_$v.so; fi
  echo "== $v"; REPS=20 python tools/prof_ops.py $SHAPES 2>&1 | grep -v Warn
done
