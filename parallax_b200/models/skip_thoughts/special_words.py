"""Special word ids (`examples/skip_thoughts/data/special_words.py`)."""
EOS, EOS_ID = "<eos>", 0
UNK, UNK_ID = "<unk>", 1
