"""Import the REAL reference modules (``/root/reference/trl``) on a box that lacks their third-party dependencies:
every module under the roots below is fabricated on demand, and every attribute of such a module is a do-nothing class.
Enough for ``import trl.trainer.ppo_trainer`` & co. to execute their module bodies, which is all ``patch_trl`` needs.
TEST INFRASTRUCTURE ONLY."""
import importlib.abc
import importlib.machinery
import sys
import types

ROOTS = ("accelerate", "peft", "unsloth", "unsloth_zoo", "deepspeed", "wandb", "vllm", "vllm_ascend", "torch_npu",
         "sentence_transformers", "llm_blender", "mergekit", "liger_kernel", "diffusers", "rich")


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def install():
    # transformers decides what is installed with find_spec: let it look before the fabricated modules exist
    import transformers  # noqa: F401
    import transformers.integrations  # noqa: F401
    import transformers.trainer  # noqa: F401
    import transformers.trainer_callback  # noqa: F401
    import transformers.training_args  # noqa: F401
    from transformers import GenerationConfig, PreTrainedModel, Trainer, TrainerCallback  # noqa: F401
    sys.meta_path.insert(0, _Finder())
