"""Drop-in replacements for the reference's `model` package (EgoVLPv2/model/*.py): same class names, constructor
arguments, attribute tree, forward signatures and state_dict keys; the arithmetic runs in libegovlp_b200.so."""
