/*
 * cr_json.h — a small JSON reader for c-ray scene files.
 * The reference uses the vendored cJSON 1.7.14 (src/libraries/cJSON.c); the behaviours its loader relies on are
 * kept: object lookup is CASE-INSENSITIVE and returns the first match (cJSON_GetObjectItem), numbers are parsed
 * with strtod into a double plus a saturating int (valuedouble / valueint), arrays keep order.
 */
#pragma once
#include <stdbool.h>
#include <stddef.h>

enum crj_type { CRJ_NULL, CRJ_FALSE, CRJ_TRUE, CRJ_NUMBER, CRJ_STRING, CRJ_ARRAY, CRJ_OBJECT };

struct crj {
	enum crj_type type;
	char *key;              /* member name when the parent is an object */
	char *str;              /* CRJ_STRING */
	double num;             /* CRJ_NUMBER: valuedouble */
	int inum;               /* CRJ_NUMBER: valueint (saturating cast) */
	struct crj *child;      /* first element / member */
	struct crj *next;       /* next sibling */
};

struct crj *crj_parse(const char *text);              /* NULL on syntax error */
void crj_free(struct crj *j);
const struct crj *crj_get(const struct crj *obj, const char *key);   /* case-insensitive, first match, NULL-safe */
const struct crj *crj_at(const struct crj *arr, int index);
int crj_size(const struct crj *arr);
static inline bool crj_is_number(const struct crj *j) { return j && j->type == CRJ_NUMBER; }
static inline bool crj_is_string(const struct crj *j) { return j && j->type == CRJ_STRING && j->str; }
static inline bool crj_is_array(const struct crj *j) { return j && j->type == CRJ_ARRAY; }
static inline bool crj_is_object(const struct crj *j) { return j && j->type == CRJ_OBJECT; }
static inline bool crj_is_bool(const struct crj *j) { return j && (j->type == CRJ_TRUE || j->type == CRJ_FALSE); }
static inline bool crj_is_true(const struct crj *j) { return j && j->type == CRJ_TRUE; }
