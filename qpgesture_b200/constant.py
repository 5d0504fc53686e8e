"""Matcher constants.

Same names and values as the reference's
codebook/Speech2GestureMatching/constant.py:26-56 (only the ones the
CodeKNN path reads), plus the derived window geometry used by the packed
database.
"""

NUM_AUDIO_FEAT_FRAMES = 6   # constant.py:26  taps per stacked audio feature
NUM_MFCC_FEAT = 13          # constant.py:32
NUM_JOINTS = 135            # constant.py:35  15 joints x 9 rotation-matrix entries
STEP_SZ = 4                 # constant.py:36  codes emitted per query step
FRAME_INTERVAL = 4          # constant.py:38  (WavLM taps use FRAME_INTERVAL-2)

num_frames = 240            # constant.py:54
num_frames_code = 30        # constant.py:55
codebook_size = 512         # constant.py:56

WAVVQ_FRAMES = 398          # literal in GestureKNN.py:436,438,632
SEED_VALUE = 123456         # GestureKNN.py:19

# Derived: every database sequence contributes this many candidate windows
# (GestureKNN.py:672 `while k < n_db_frm - STEP_SZ*step_sz`, :713 range(0, 240-32, 8)).
WINDOWS_PER_SEQ = num_frames_code - STEP_SZ   # 26
STEPS_PER_SEGMENT = 8       # GestureKNN.py:528,659: i += STEP_SZ*step_sz until len(clip)
EMPTY_DIST = 1e3            # GestureKNN.py:668,709 sentinel for a start-code with no window
