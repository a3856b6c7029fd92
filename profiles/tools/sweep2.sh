for parts in 1 2 3 4 8 12; do
  SPEEDY_B200_WRITE_PARTS=$parts python profiles/tools/step_time.py 2>&1 | tail -1
  SPEEDY_K4_CHAIN=0 SPEEDY_B200_WRITE_PARTS=$parts python profiles/tools/step_time.py 2>&1 | tail -1
done
