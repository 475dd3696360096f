for t in 256 320 384 448 512; do echo threads $t; GTA_B200_LPF_THREADS=$t GTA_B200_LPF_SLOTS=1152 python scratch/lpf_time.py 100000 2>&1 | tail -2; done
echo "fused:"; GTA_B200_LPF=0 python scratch/lpf_time.py 20000 2>&1 | tail -2
