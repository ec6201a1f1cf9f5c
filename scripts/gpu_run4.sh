for cfg in "patch_dbg=0" "patch_dbg=1" "patch_dbg=2" "patch_dbg=0 patch_tiles=0" "patch_dbg=1 patch_tiles=0" "patch_dbg=2 patch_tiles=0"; do
  echo "== $cfg"; timeout 300 python scripts/patch_timeline.py 707 $cfg | grep -E "alive"
done
