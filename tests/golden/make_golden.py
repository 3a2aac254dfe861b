"""Regenerate tests/golden/ from the reference checkout (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

Copies the reference's own golden fixtures for the simulation path:
  * tests/fixtures/{orbium,orbium-scutium,aquarium}-test.yaml      (test configs, data not code)
  * tests/fixtures/*_last_frame*.p  → <name>_last_frame.npy        (float32 last frames asserted by
    tests/test_pipeline.py:34-35,53-54,72-73,91-92,110-111,129-130)
  * conf/species/2d/1c-1k/orbium.yaml                              (BASELINE config A)
and, using the oracle only to decode them, stores decoded initial cells so the decoders themselves
are pinned by a checksum test.  /root/reference is NOT available on the GPU box, hence the copy.
"""
import os
import pickle
import shutil
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
fx = os.path.join(ref, 'tests', 'fixtures')

pairs = {
    'orbium-test': 'orbium-test_last_frame.p',
    'orbium-scutium-test': 'orbium-scutium-test_last_frame2.p',
    'aquarium-test': 'aquarium-test_last_frame.p',
}
for name, frame in pairs.items():
    shutil.copyfile(os.path.join(fx, name + '.yaml'), os.path.join(here, name + '.yaml'))
    with open(os.path.join(fx, frame), 'rb') as f:
        arr = np.asarray(pickle.load(f), dtype=np.float32)
    np.save(os.path.join(here, name + '_last_frame.npy'), arr)
    print(name, arr.shape, arr.dtype, float(arr.sum()))

shutil.copyfile(os.path.join(ref, 'conf', 'species', '2d', '1c-1k', 'orbium.yaml'), os.path.join(here, 'orbium.yaml'))
print('done')
