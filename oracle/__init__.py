"""CPU oracle for the SpeechSplit feature front end (TEST INFRASTRUCTURE ONLY).

Everything under ``oracle/`` is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  The product package ``speechsplit_b200`` never does.

Parity status:
  * stages a0-a3, a5, a7-a9 (filtfilt, dither, pySTFT, mel-dB, normalisation,
    quantisation): PINNED - the restatement is checked against the reference's own
    ``utils.py`` imported in the build container (tests/golden/make_golden.py).
  * InterpLnr (oracle/interp_lnr.py): PINNED - bit-identical to the reference's own ``model.InterpLnr``
    run in the build container on captured random draws (tests/golden/interp_lnr.npz).
  * training collator and manifest (oracle/collate_ref.py): PINNED - bit-identical to the batch of the
    reference's own ``data_loader.MyCollator`` and the ``train.pkl`` of its ``make_metadata.py``
    (tests/golden/collate.npz).
  * stage a4 (librosa.filters.mel) and stage a6 (pysptk.sptk.rapt -> SPTK/Snack get_f0):
    PARITY UNPINNED - neither package (nor its source) exists in the build container, so
    these are restatements of the published algorithms, anchored on the reference's call
    sites (make_spect_f0.py:15, :64) and two independent cross-checks for the mel (torchaudio,
    transformers.audio_utils).
"""
