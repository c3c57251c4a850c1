# A/B timing of study builds of libbmpc on ONE box (box-to-box the same binary differs by ~2 %): every
# tools/studies/build/ab_*.so runs nscale.py at N = 4096, interleaved, ROUNDS times.
for r in $(seq ${ROUNDS:-2}); do
  for f in tools/studies/build/ab_*.so; do
    echo "$(basename $f): $(BMPC_LIB=$PWD/$f python tools/studies/nscale.py 4096 2>/dev/null | head -1 | cut -c1-75)"
  done
done
