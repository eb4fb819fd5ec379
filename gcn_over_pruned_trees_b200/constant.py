"""Id spaces the hot path depends on.

The kernels hard-code three facts about the reference's id space
(/root/reference/utils/constant.py:12-17,29,35):

* a reverse edge carries ``deprel + 42``         (DEPREL_FORWARD_BOUND)
* a self loop carries id ``84``                  (SELF_LOOP_INDEX)
* masked max pooling fills with ``-1e12``        (INFINITY_NUMBER)

The embedding tables are sized from the *counts* of the POS / NER / DEPREL id maps
(47 / 15 / 85, /root/reference/model/gcn.py:45-57), so only the counts are needed to keep the
checkpoint ``state_dict`` layout; the string->id maps below are rebuilt from ordered name
lists so that data prepared with the reference's loader indexes the same rows.
"""

PAD_TOKEN = '<PAD>'
PAD_ID = 0
UNK_TOKEN = '<UNK>'
UNK_ID = 1
VOCAB_PREFIX = [PAD_TOKEN, UNK_TOKEN]
EMB_INIT_RANGE = 1.0

SELF_LOOP = 'self_loop'
DEPREL_FORWARD_BOUND = 42
DEPREL_REVERSE_BOUND = 84
SELF_LOOP_INDEX = 84
INFINITY_NUMBER = 1e12


def _enumerate(names):
    return {name: i for i, name in enumerate(names)}


_NER_NAMES = VOCAB_PREFIX + ['O', 'PERSON', 'ORGANIZATION', 'LOCATION', 'DATE', 'NUMBER', 'MISC', 'DURATION',
                             'MONEY', 'PERCENT', 'ORDINAL', 'TIME', 'SET']

_POS_NAMES = VOCAB_PREFIX + ['NNP', 'NN', 'IN', 'DT', ',', 'JJ', 'NNS', 'VBD', 'CD', 'CC', '.', 'RB', 'VBN', 'PRP',
                             'TO', 'VB', 'VBG', 'VBZ', 'PRP$', ':', 'POS', "''", '``', '-RRB-', '-LRB-', 'VBP',
                             'MD', 'NNPS', 'WP', 'WDT', 'WRB', 'RP', 'JJR', 'JJS', '$', 'FW', 'RBR', 'SYM', 'EX',
                             'RBS', 'WP$', 'PDT', 'LS', 'UH', '#']

_DEPREL_FORWARD = VOCAB_PREFIX + ['punct', 'compound', 'case', 'nmod', 'det', 'nsubj', 'amod', 'conj', 'dobj',
                                  'ROOT', 'cc', 'nmod:poss', 'mark', 'advmod', 'appos', 'nummod', 'dep', 'ccomp',
                                  'aux', 'advcl', 'acl:relcl', 'xcomp', 'cop', 'acl', 'auxpass', 'nsubjpass',
                                  'nmod:tmod', 'neg', 'compound:prt', 'mwe', 'parataxis', 'root', 'nmod:npmod',
                                  'expl', 'csubj', 'cc:preconj', 'iobj', 'det:predet', 'discourse', 'csubjpass']

_TACRED_LABELS = ['no_relation', 'per:title', 'org:top_members/employees', 'per:employee_of',
                  'org:alternate_names', 'org:country_of_headquarters', 'per:countries_of_residence',
                  'org:city_of_headquarters', 'per:cities_of_residence', 'per:age',
                  'per:stateorprovinces_of_residence', 'per:origin', 'org:subsidiaries', 'org:parents',
                  'per:spouse', 'org:stateorprovince_of_headquarters', 'per:children', 'per:other_family',
                  'per:alternate_names', 'org:members', 'per:siblings', 'per:schools_attended', 'per:parents',
                  'per:date_of_death', 'org:member_of', 'org:founded_by', 'org:website', 'per:cause_of_death',
                  'org:political/religious_affiliation', 'org:founded', 'per:city_of_death', 'org:shareholders',
                  'org:number_of_employees/members', 'per:date_of_birth', 'per:city_of_birth', 'per:charges',
                  'per:stateorprovince_of_death', 'per:religion', 'per:stateorprovince_of_birth',
                  'per:country_of_birth', 'org:dissolved', 'per:country_of_death']

NER_TO_ID = _enumerate(_NER_NAMES)
POS_TO_ID = _enumerate(_POS_NAMES)
# forward ids 0..41, reverse ids 42..83 (= forward + 42), self loop 84
DEPREL_TO_ID = _enumerate(_DEPREL_FORWARD + [n + '_reverse' for n in _DEPREL_FORWARD] + [SELF_LOOP])
LABEL_TO_ID = _enumerate(_TACRED_LABELS)
NEGATIVE_LABEL = 'no_relation'

NUM_POS = len(POS_TO_ID)        # 47
NUM_NER = len(NER_TO_ID)        # 15
NUM_DEPREL = len(DEPREL_TO_ID)  # 85
ROOT_DEPREL_ID = DEPREL_TO_ID['ROOT']  # 11

assert NUM_POS == 47 and NUM_NER == 15 and NUM_DEPREL == 85
assert len(_DEPREL_FORWARD) == DEPREL_FORWARD_BOUND and DEPREL_TO_ID[SELF_LOOP] == SELF_LOOP_INDEX
assert len(LABEL_TO_ID) == 42
