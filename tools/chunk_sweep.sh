cd /root/repo
run() { python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" | python -c "import json,sys; d=json.load(sys.stdin); print('%.4g  %.4f ms/step  frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"; }
for mc in 16 22 32 64; do export LH_MIN_CHUNK=$mc; echo "min chunk $mc"; echo -n "  coupled 64: "; run; echo -n "  general 64: "; run --general-vg; echo -n "  richards 100: "; run --model richards --nlayer 100 --ncol 655360; echo -n "  coupled 131072 cols: "; run --ncol 131072; done
