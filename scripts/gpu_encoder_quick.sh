# encoder parity tests + the scorer figures that matter when tuning encode_kernel
IIV_RANDOM_CASES=${CASES:-40} timeout 400 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_encoder_random.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --scorer-only 2>/dev/null | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())['scorer']
t = d['encoded_trace_clip0']; s = d['single_clip_trace']
print('148 clips %.0f fps (slowest %d cyc; clip0 A %d loop %d wait_rows %d wait_mt %d)' % (
    d['encoded_frames_per_s'], t['slowest_clip_sm_cycles'], t['cycles_score_heapify'],
    t['cycles_opcode_loop'], t['cycles_wait_rows'], t['cycles_wait_mt_applier']))
print('single %.1f fps (A %d loop %d wait_rows %d wait_mt %d)  hgr60 %.1f fps  facade %.0f  movie %.0f' % (
    d['single_clip_frames_per_s'], s['cycles_score_heapify'], s['cycles_opcode_loop'],
    s['cycles_wait_rows'], s['cycles_wait_mt_applier'], d['hgr_60_frame_clip_frames_per_s'],
    d.get('facade_frames_per_s', 0), d.get('movie_frames_per_s', 0)))"
