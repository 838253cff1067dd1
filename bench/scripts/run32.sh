# knob sweep of the batched-affine kernel at shard size (2^21 points, c = 20)
for V in "BMPC_AFF_WAVES=1" "BMPC_AFF_WAVES=3" "BMPC_AFF_WAVES=1 BMPC_AFF_BLOCKDIM=128" "BMPC_AFF_KSEL=128" "BMPC_AFF_GMAX=2" "BMPC_AFF_GMAX=3" "BMPC_AFF_WAVES=4"; do
echo "[$V] $(env $V python bench/shard_window.py 21 20 2>&1 | tail -1)"
done
