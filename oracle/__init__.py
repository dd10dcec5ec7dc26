"""TEST INFRASTRUCTURE ONLY.

`oracle/` is the CPU restatement of the reference's post-processing hot path (plus build
recipes for the parts of the reference itself that compile here, see build_ref.py). It is the
checker for tests/, `__graft_entry__.smoke()` and the CPU-baseline legs of bench.py. Nothing
under `pytorchocr_b200/` may import it: the product path has no CPU fallback.

Parity pinning status: the reference ships NO tests, fixtures or golden vectors for this path
(SURVEY.md §4). The oracle is pinned against outputs of the reference itself run in the
authoring container: its compiled Cython modules (oracle/_ref/pse, pa), its vendored Clipper
(oracle/_ref/libclipper_ref.so) and its unmodified Python operator classes
(tests/golden/make_golden.py -> tests/golden/*.npz). The DB C++ module cannot be built here
(no OpenCV C++), so for DB the oracle is a line-by-line restatement over cv2-python 4.13:
parity for DB is pinned only through that restatement ("parity unpinned" upstream).
The text-line crop oracle (crop_oracle.py) is the reference's own sequence of cv2 calls; it is pinned against
the reference's `sort_boxes` / `get_part_img` imported in place (tests/golden/reference_crops.npz) and, bit for
bit, against a first-principles restatement of what those cv2 calls compute (tests/test_oracle_crop.py).
"""
