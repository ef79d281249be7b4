for wl in iterative_480x640_1Mev iterative_480x640_1Mev_edges iterative_128x128_b8_f4 iterative_480x640_250kev; do
for mb in 12 16 24 32 48; do echo "band_mb=$mb"; TEF_BAND_BYTES=$((mb<<20)) python scripts/kernel_times.py --workload $wl 2>&1 | tail -1 | cut -c1-260; done; done
