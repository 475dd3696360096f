python scratch/lpf_time.py 100000 2>&1 | tail -3
echo "== cfg3 fused"; GTA_B200_LPF=0 python scratch/lpf_vs_cta.py 3 16384 2>&1 | tail -1
echo "== cfg5"; GTA_B200_LPF=1 python scratch/lpf_vs_cta.py 5 50000 2>&1 | tail -1; GTA_B200_LPF=0 python scratch/lpf_vs_cta.py 5 10000 2>&1 | tail -1
echo "== cfg1"; GTA_B200_LPF=1 python scratch/lpf_vs_cta.py 1 100000 2>&1 | tail -1
