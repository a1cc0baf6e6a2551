for tw in 74 148 296 600; do for lm in 64 128; do
echo "TILE_WAVE=$tw L_MIN=$lm"; SPLLT_B200_TILE_WAVE=$tw SPLLT_B200_TILE_L_MIN=$lm timeout 200 python scratch/g2.py 64 2>&1 | grep -E "  factor|  profile"
done; done
