"""fragnet_b200: B200-native (sm_100a) implementation of FragNet's GAT2 message-passing hot path.

Where things are (reference lines in the module docstrings; design and measurements in DESIGN.md):

* ``model.gat.gat2`` / ``gat2_lite`` / ``gat2_pretrain`` / ``pretrain_heads`` -- drop-in ``FragNetLayerA``, ``FragNet``,
  ``FragNetFineTune``, ``FragNetPreTrain``, ``PretrainTask``, ``FTHead*`` (re-exported under the reference's own module
  paths by the top-level ``fragnet`` package);
* ``vizualize.model`` -- attention-returning ``FragNetViz`` / ``FragNetFineTuneViz`` / ``FragNetPreTrainViz``;
  ``vizualize.attribution`` -- every atom / bond / fragment-link mask of a molecule in one forward;
* ``train.pretrain_utils.Trainer``, ``train.utils.TrainerFineTune`` / ``EarlyStopping`` / ``test_fn`` -- the
  reference's loops; ``train.fused.FusedPretrainStep`` -- one library call per pretraining step;
* ``dataset.data`` -- ``collate_fn`` / ``collate_fn_pt``; ``dataset.prefetch.DevicePrefetcher`` -- overlapped staging of
  host batches; ``dataset.arena.MoleculeArena`` / ``ArenaLoader`` -- dataset resident in HBM, batches assembled on the
  device; ``screen.screen`` -- pipelined inference screening;
* ``dist`` -- flat-gradient NCCL all-reduce; ``config`` -- precision switch (fp32 parity mode / tf32 tensor cores);
* ``ops`` / ``autograd`` / ``_abi`` -- ctypes plumbing over ``include/fragnet_b200.h`` (``csrc/*.cu``).

There is no CPU path: every op needs the in-tree CUDA library and fails loudly without it.
"""
__version__ = "0.1.0"
