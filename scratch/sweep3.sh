export GTA_B200_LPF_SLOTS=1152
for t in 384 448 512; do echo threads $t; GTA_B200_LPF_THREADS=$t python scratch/lpf_time.py 100000 2>&1 | tail -2 | head -1; done
export GTA_B200_LPF_THREADS=448
echo "== cfg3 fused"; GTA_B200_LPF=0 python scratch/lpf_vs_cta.py 3 16384 2>&1 | tail -1
echo "== cfg5"; GTA_B200_LPF=1 python scratch/lpf_vs_cta.py 5 50000 2>&1 | tail -1; GTA_B200_LPF_THREADS=384 GTA_B200_LPF=1 python scratch/lpf_vs_cta.py 5 50000 2>&1 | tail -1; GTA_B200_LPF=0 python scratch/lpf_vs_cta.py 5 10000 2>&1 | tail -1
