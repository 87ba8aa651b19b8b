# same-box A/B of attention64 variants: bash tools/a64_ab.sh name1 name2 ...  (variants/libwfl_<name>.so; "main" = in-tree)
for i in 1 2 3; do
  for v in main "$@"; do
    echo -n "$v  "
    if [ "$v" = main ]; then python tools/prof_ops.py attn64; else WFL_LIB=variants/libwfl_$v.so python tools/prof_ops.py attn64; fi
  done
done
for v in main "$@"; do
  echo -n "$v B64 H12 "
  if [ "$v" = main ]; then PROF_B=64 PROF_D=768 python tools/prof_ops.py attn64; else WFL_LIB=variants/libwfl_$v.so PROF_B=64 PROF_D=768 python tools/prof_ops.py attn64; fi
done
