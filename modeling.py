"""Drop-in module name: the reference's callers do `from modeling import MM_LLMs, MM_LLMs_Config`
(run_clm_llms.py:77, llm_trainer.py:23).  Everything lives in macaw-llm_b200/modeling.py."""
from macaw_llm_b200.modeling import LlamaForCausalLM, LlamaModel, MM_LLMs, MM_LLMs_Config, MM_LLMsConfig  # noqa: F401
