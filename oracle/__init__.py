"""CPU oracle for the DEFT per-frame dense hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deft_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or as the timed CPU baseline -- never as the thing shipped.

Contents
--------
``spec``      parameter inventory of DLASeg(dla34)+AFE+DecoderRNN (names/shapes of
              the reference's ``state_dict``) and the seeded synthetic initialiser
              of SURVEY.md section 8(d).
``restate``   plain fp32 PyTorch-CPU restatement of the reference algorithm, written
              functionally over a ``state_dict`` (each function cites the reference
              ``file:line`` it follows).
``ref_loader`` imports the *real* reference modules from ``/root/reference`` (only
              available in the build container) to pin the restatement and to
              generate ``tests/golden``.

Parity pin: the reference ships no tests/golden vectors, so the
restatement is pinned against outputs of the reference's own modules executed
in the build container (``oracle/gen_golden.py`` -> ``tests/golden/*.pt``).
The DCNv2 arithmetic lives in the un-vendored CharlesShang/DCNv2 extension
(unpinned ``git clone``, reference ``README.md:72-78``); its executable stand-in
is ``torchvision.ops.deform_conv2d`` (same offset/mask convention), so the DCN
row is "pinned to torchvision's deform_conv2d, upstream DCNv2 itself unpinned".
"""
