# Same-box A/B of the block pose kernel's chain walk: lane = body (product) against lane = joint
# (python tools/build_variant.py 3d-human-body-reconstruction_b200/libsmplk_lanejoint.so -DSMPLK_POSE_FWD_LANE_BODY=0), alternating
V=$PWD/3d-human-body-reconstruction_b200/libsmplk_lanejoint.so
for i in 1 2; do
  echo "== lane = body (product)"; python tools/pdl_ab.py 4096 16384
  echo "== lane = joint (variant)"; SMPLK_LIB=$V python tools/pdl_ab.py 4096 16384
done
