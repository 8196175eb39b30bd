"""fragnet_b200: B200-native (sm_100a) implementation of FragNet's GAT2 message-passing hot path.

Where things are (reference lines in the module docstrings; design and measurements in DESIGN.md):

* ``model.gat.gat2`` / ``gat2_lite`` / ``gat2_edge`` / ``gat2_pretrain`` / ``pretrain_heads`` -- drop-in ``FragNetLayerA``, ``FragNet``,
  ``FragNetFineTune``, ``FragNetPreTrain``, ``PretrainTask``, ``FTHead*`` (re-exported under the reference's own module
  paths by the top-level ``fragnet`` package);
* ``vizualize.model`` -- attention-returning ``FragNetViz`` / ``FragNetFineTuneViz`` / ``FragNetPreTrainViz``;
  ``vizualize.attribution`` -- every atom / bond / fragment-link mask of a molecule in one forward;
* ``train.pretrain_utils.Trainer``, ``train.utils.TrainerFineTune`` / ``EarlyStopping`` / ``test_fn`` -- the
  reference's loops; ``train.fused.FusedPretrainStep`` -- one library call per pretraining step (the next batch's
  collate queued underneath it, gradient exchange + Adam as one kernel over NVLink peer memory under torch.distributed);
* ``dataset.data`` -- ``collate_fn`` / ``collate_fn_pt`` and their compact (uint8 / int32) and packed (one buffer) wire
  formats; ``dataset.prefetch.DevicePrefetcher`` -- overlapped staging of host batches, widened on the device;
  ``dataset.arena.MoleculeArena`` / ``ArenaLoader`` -- dataset resident in HBM, batches assembled on the device
  underneath the step in flight; ``screen.screen`` -- pipelined inference screening;
* ``dist`` -- flat-gradient NCCL all-reduce; ``config`` -- precision switch (``fp32`` = 3xTF32 split on the tensor cores,
  the 1e-5 parity mode and default / ``tf32`` / ``fp32_simt``);
* ``ops`` / ``autograd`` / ``_abi`` -- ctypes plumbing over ``include/fragnet_b200.h`` (``csrc/*.cu``); ``_compat`` --
  overlay of the top-level ``fragnet`` shim on a reference checkout for everything outside the hot path.

There is no CPU path: every op needs the in-tree CUDA library and fails loudly without it.
"""
__version__ = "0.1.0"
