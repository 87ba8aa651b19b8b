# same-box A/B of the first-generation attention kernel's softmax forms: WFL_ATTN_V1_QUAD=1 (row-quad) vs 0 (column split)
for q in 1 0 1 0; do
  export WFL_ATTN_V1_QUAD=$q
  echo -n "quad=$q hd256 B32: "; python tools/prof_ops.py attn256
  echo -n "quad=$q hd384 B64 d768: "; PROF_B=64 PROF_D=768 python tools/prof_ops.py attn384
  echo -n "quad=$q bias B16 T799 H16: "; PROF_B=16 PROF_T=799 PROF_D=1024 python tools/prof_ops.py attn64b
  echo -n "quad=$q bias B32 T1499 H12: "; PROF_B=32 PROF_T=1499 PROF_D=768 python tools/prof_ops.py attn64b
  echo -n "quad=$q v1 hd64 plain: "; WFL_ATTN64=v1 python tools/prof_ops.py attn64
done
